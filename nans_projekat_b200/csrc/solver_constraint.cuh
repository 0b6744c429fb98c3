// solver_constraint.cuh — the arithmetic of one Constraint (code/nans.cpp:1021-1329), in the two halves the solver
// kernels use.  Pure device functions over DeviceWorld: tests/test_solver_host.py compiles this header for the
// host, runs the reference's sequential sweep with it and compares with the oracle bit for bit.
#pragma once
#include "nans_math.cuh"
#include "solver_accum.cuh"
#include "world.cuh"

namespace nans {

// Constraint, code/nans.cpp:1021-1329 — bug-compatible: minus sign on body B's
// angular JMJ term, cross(W, N) instead of cross(W, R), un-normalised T1, 70 iterations over
// constants of which only the last delta is applied, friction bound evaluated in fp64.
//
// Split in two so that the dependency-critical part is as short as possible:
//   contact_prep_kernel  everything that does not read velocities (N, T1, T2, the six R x axis
//                        vectors, the three effective masses, the Baumgarte term, inverse masses) —
//                        embarrassingly parallel, one 160-byte record per contact;
//   apply_prepared       what must wait for the predecessors: relative velocities, the 70-iteration
//                        accumulation, the impulses.  One contiguous record load + four body rows.
// The arithmetic (operations, order, roundings) is exactly that of the single function.
constexpr int kRecQuads = 11;   // float4 per contact record

// q[0..8] of contact c (everything that does not read velocities); returns the body rows
__device__ __forceinline__ void constraint_prepare(const DeviceWorld &w, int c, float dt, float4 (&q)[kRecQuads], int &ia, int &ib)
{
    const float4 cpa = w.c_pa[c], cpb = w.c_pb[c];      // w lanes carry the body rows
    ia = __float_as_int(cpa.w); ib = __float_as_int(cpb.w);
    const vec3 posA = V3(w.pos[ia]);
    const float invM1 = w.vel[ia].w, invI1 = w.angvel[ia].w;
    vec3 posB;
    float invM2, invI2;
    if (ib >= 0) {
        posB = V3(w.pos[ib]);
        invM2 = w.vel[ib].w; invI2 = w.angvel[ib].w;
    } else {
        const int k = -ib - 1;                 // the Floor (:1065-1077)
        const float4 sp = w.st_pos[k], sa = w.st_ang[k];
        posB = V3(sp);
        invM2 = sp.w; invI2 = sa.w;
    }
    vec3 N = normalize(V3(w.c_n[c]));
    if (equal(N, V3(0.f, 0.f, 0.f))) N = normalize(posB - posA);   // :1115-1119
    const vec3 R1 = V3(cpa) - posA;
    const vec3 R2 = V3(cpb) - posB;
    vec3 T1;
    if (N.x >= 0.57735f) T1 = V3(N.y, -N.x, 0.0f); else T1 = V3(0.0f, N.z, -N.y);
    const vec3 T2 = cross(N, T1);                                    // T1 is NOT normalised (:1133)
    const float depth = dot((posA + R1) - (posB + R2), N);
    const vec3 RN1 = cross(R1, N), RN2 = cross(R2, N);
    float JMJn = fadd(invM1, invM2);
    JMJn = fadd(JMJn, fsub(fmul(invI1, dot(RN1, RN1)), fmul(invI2, dot(-RN2, -RN2))));
    JMJn = frcp(JMJn);
    const vec3 R1T1 = cross(R1, T1), R2T1 = cross(R2, T1), R1T2 = cross(R1, T2), R2T2 = cross(R2, T2);
    float JMJt1 = fadd(invM1, invM2);
    JMJt1 = fadd(JMJt1, fsub(fmul(invI1, dot(R1T1, R1T1)), fmul(invI2, dot(-R2T1, -R2T1))));
    JMJt1 = frcp(JMJt1);
    float JMJt2 = fadd(invM1, invM2);
    JMJt2 = fadd(JMJt2, fsub(fmul(invI1, dot(R1T2, R1T2)), fmul(invI2, dot(-R2T2, -R2T2))));
    JMJt2 = frcp(JMJt2);
    const float Bd = fmul(fdiv(-0.3f, dt), depth);                  // (-Beta / dt) * Depth (:1156)
    q[0] = make_float4(N.x, N.y, N.z, JMJn);
    q[1] = make_float4(T1.x, T1.y, T1.z, JMJt1);
    q[2] = make_float4(T2.x, T2.y, T2.z, JMJt2);
    q[3] = make_float4(RN1.x, RN1.y, RN1.z, invM1);
    q[4] = make_float4(RN2.x, RN2.y, RN2.z, invM2);
    q[5] = make_float4(R1T1.x, R1T1.y, R1T1.z, invI1);
    q[6] = make_float4(R2T1.x, R2T1.y, R2T1.z, invI2);
    q[7] = make_float4(R1T2.x, R1T2.y, R1T2.z, Bd);
    q[8] = make_float4(R2T2.x, R2T2.y, R2T2.z, 0.f);
}

// The velocity-dependent part of Constraint (code/nans.cpp:1158-1328) on the prepared record q[0..8]:
// relative velocities, the 70-iteration accumulation, the impulses.  has_b == false: body B is the
// Floor (V = W = 0, never written, :1278-1289).
__device__ __forceinline__ void constraint_apply(const float4 (&q)[kRecQuads], vec3 &V1, vec3 &W1, vec3 &V2, vec3 &W2,
                                                 bool has_b)
{
    const vec3 N = V3(q[0]), T1 = V3(q[1]), T2 = V3(q[2]);
    const float JMJn = q[0].w, JMJt1 = q[1].w, JMJt2 = q[2].w;
    const vec3 RN1 = V3(q[3]), RN2 = V3(q[4]), R1T1 = V3(q[5]), R2T1 = V3(q[6]), R1T2 = V3(q[7]), R2T2 = V3(q[8]);
    const float invM1 = q[3].w, invM2 = q[4].w, invI1 = q[5].w, invI2 = q[6].w, Bd = q[7].w;
    const vec3 dVn = ((V1 + cross(W1, N)) - V2) - cross(W2, N);
    const float JdVn = dot(dVn, N);
    const float B = fadd(Bd, fmul(0.1f, JdVn));                        // + Cr * JdVn
    const vec3 dVt1 = ((V1 + cross(W1, T1)) - V2) - cross(W2, T1);
    const float JdVt1 = dot(dVt1, T1);
    const vec3 dVt2 = ((V1 + cross(W1, T2)) - V2) - cross(W2, T2);
    const float JdVt2 = dot(dVt2, T2);

    // :1176-1227 — the accumulators start at zero (SolveConstraints works on a copy, :1545)
    const float lambdaN = fmul(fadd(-JdVn, B), JMJn);
    const float lambdaT1 = fmul(-JdVt1, JMJt1);
    const float lambdaT2 = fmul(-JdVt2, JMJt2);
    // the 70-iteration accumulation (solver_accum.cuh); NaN increments poison the sums: literal compares
    AccumDeltas acc;
    if (lambdaN == lambdaN && lambdaT1 == lambdaT1 && lambdaT2 == lambdaT2) {
#if NANS_ACCUM_FAST
        acc = accumulate_fast(lambdaN, lambdaT1, lambdaT2);
#else
        acc = accumulate_pipelined(lambdaN, lambdaT1, lambdaT2);
#endif
    } else {
        acc = accumulate_literal(lambdaN, lambdaT1, lambdaT2);
    }
    const float DLN = acc.DLN, DLT1 = acc.DLT1, DLT2 = acc.DLT2;
    const vec3 LI = N * DLN, LIT1 = T1 * DLT1, LIT2 = T2 * DLT2;
    const vec3 AI1 = RN1 * DLN, AI2 = RN2 * DLN;
    const vec3 AI1T1 = R1T1 * DLT1, AI2T1 = R2T1 * DLT1;
    const vec3 AI1T2 = R1T2 * DLT2, AI2T2 = R2T2 * DLT2;
    // :1229-1328 — normal, then T1, then T2; a != b so register accumulation equals the
    // reference's read-modify-write sequence
    V1 = V1 + invM1 * LI;     W1 = W1 + invI1 * AI1;
    V1 = V1 + invM1 * LIT1;   W1 = W1 + invI1 * AI1T1;
    V1 = V1 + invM1 * LIT2;   W1 = W1 + invI1 * AI1T2;
    if (has_b) {
        V2 = V2 - invM2 * LI;     W2 = W2 - invI2 * AI2;
        V2 = V2 - invM2 * LIT1;   W2 = W2 - invI2 * AI2T1;
        V2 = V2 - invM2 * LIT2;   W2 = W2 - invI2 * AI2T2;
    }
}

}  // namespace nans
