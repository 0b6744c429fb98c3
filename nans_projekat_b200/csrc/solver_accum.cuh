// solver_accum.cuh — the 70-iteration impulse accumulation of Constraint (code/nans.cpp:1176-1227).
//
// The reference iterates 70 times over CONSTANT increments (velocities are not re-read) and applies only
// the LAST iteration's deltas; the rounding of the running fp32 sums is observable, so the sums have to
// be run.  Per iteration k:
//     N_k = max(fl(N_{k-1} + LN), 0)
//     M_k = (float)((sqrt(2.0) * (double)0.1f) * (double)N_k)            friction bound, evaluated in fp64
//     T_k = min(max(fl(T_{k-1} + LT), -M_k), M_k)                        for each of the two tangents
// and the result is DLN = N_70 - N_69, DLT = T_70 - T_69.
//
// accumulate_literal   the reference's statements one by one (used when an increment is NaN, and as the
//                      fallback and the test oracle of the fast form).
// accumulate_pipelined the same with min/max and the fp64 bound software-pipelined ahead of the tangent chains
//                      (round 1's kernel form): ~43 cycles per iteration, bound by the two fp64 <-> fp32
//                      conversions (XU pipe) and the FADD -> FMNMX -> FMNMX chain of a tangent sum.
// accumulate_fast      same bits from THREE PLAIN FADD CHAINS (round 2).  With LN > 0 the normal sum never
//                      clamps, and a tangent is, for all 70 iterations, in one of two regimes that one compare
//                      against LN decides:
//                        free      |LT| <= (1 - 1e-4) c LN : the clamp never fires, T_k is the running sum of LT;
//                        saturated |LT| >= (1 + 2e-3) c LN : the clamp fires every iteration, T_k = sign(LT) M_k,
//                                                            so only M_69 and M_70 are needed (two fp64 bounds);
//                      c = sqrt(2) * 0.1.  Proof sketch: k fp32 additions of a constant stay within a relative
//                      k * 2^-24 (< 4.2e-6 for k <= 70) of k * L, and M_k within 2^-24 of c N_k, so the ratio
//                      |T_k| / M_k stays within 1.3e-5 of |LT| / (c LN) while unclamped (free), and
//                      M_(k-1) + |LT| >= M_k (1 + 4e-4) by induction when saturated; LN is required to lie in
//                      [1e-30, 1e30] so that nothing under- or overflows.  Anything else (the 0.2 % band around
//                      saturation, extreme magnitudes) runs the pipelined form.  Closed forms: LN <= 0 keeps
//                      N = M = 0 and gives all deltas +0.
// tests/test_solver_accum_host.py compiles this header for the host and compares the forms bit for bit
// over random, borderline and non-finite increments.
#pragma once
#include "nans_math.cuh"

namespace nans {

#ifndef NANS_ACCUM_FAST
#define NANS_ACCUM_FAST 1   // 0: the pipelined form for every contact (round 1)
#endif

struct AccumDeltas { float DLN, DLT1, DLT2; };

// sqrt(2) * Cf in fp64 (:1197) and the two floats that bracket it
#define NANS_KFRIC (1.4142135623730951 * (double)0.1f)
#define NANS_KFRIC_LO 0x1.21a184p-3f
#define NANS_KFRIC_HI 0x1.21a186p-3f

__device__ __forceinline__ float friction_bound(float sumN)
{
    return __double2float_rn(__dmul_rn(NANS_KFRIC, (double)sumN));
}

__device__ __forceinline__ AccumDeltas accumulate_literal(float lambdaN, float lambdaT1, float lambdaT2)
{
    float DLN = 0.f, sumN = 0.f, DLT1 = 0.f, sumT1 = 0.f, DLT2 = 0.f, sumT2 = 0.f;
#pragma unroll 1
    for (int it = 0; it < 70; ++it) {
        const float oldN = sumN;
        sumN = fadd(sumN, lambdaN);
        if (sumN < 0) sumN = 0.0f;
        DLN = fsub(sumN, oldN);
        const float maxT = friction_bound(sumN);
        const float oldT1 = sumT1;
        sumT1 = fadd(sumT1, lambdaT1);
        if (sumT1 < -maxT) sumT1 = -maxT;
        if (sumT1 > maxT) sumT1 = maxT;
        DLT1 = fsub(sumT1, oldT1);
        const float oldT2 = sumT2;
        sumT2 = fadd(sumT2, lambdaT2);
        if (sumT2 < -maxT) sumT2 = -maxT;
        if (sumT2 > maxT) sumT2 = maxT;
        DLT2 = fsub(sumT2, oldT2);
    }
    AccumDeltas r; r.DLN = DLN; r.DLT1 = DLT1; r.DLT2 = DLT2;
    return r;
}

// The previous device form, kept for A/B (NANS_ACCUM_FAST=0): min/max instead of the compares (bit-identical
// without NaN increments: a NaN BOUND leaves s alone in both forms; max(+0,-0) = +0 and min(-0,+0) = -0 are
// what the compares leave too), the normal chain and the fp64 bound software-pipelined a block ahead of the
// tangent chains.  Bound by the two fp64 conversions per iteration (XU pipe): ~43 cycles per iteration.
__device__ __forceinline__ AccumDeltas accumulate_pipelined(float lambdaN, float lambdaT1, float lambdaT2)
{
    float sumN = 0.f, sumT1 = 0.f, sumT2 = 0.f;
    constexpr int kBlk = 10;
    float bound[kBlk], bound_next[kBlk];
    float oldN = 0.f, oldT1 = 0.f, oldT2 = 0.f;
    auto normal_step = [&]() -> float {
        oldN = sumN;
        sumN = fmaxf(fadd(sumN, lambdaN), 0.0f);
        return friction_bound(sumN);
    };
#pragma unroll
    for (int j = 0; j < kBlk; ++j) bound_next[j] = normal_step();
#pragma unroll 1
    for (int blk = 0; blk < 70 / kBlk; ++blk) {
#pragma unroll
        for (int j = 0; j < kBlk; ++j) bound[j] = bound_next[j];
        const bool more = blk + 1 < 70 / kBlk;
#pragma unroll
        for (int j = 0; j < kBlk; ++j) {
            if (more) bound_next[j] = normal_step();
            const float maxT = bound[j];
            oldT1 = sumT1;
            sumT1 = fminf(fmaxf(fadd(sumT1, lambdaT1), -maxT), maxT);
            oldT2 = sumT2;
            sumT2 = fminf(fmaxf(fadd(sumT2, lambdaT2), -maxT), maxT);
        }
    }
    AccumDeltas r;
    r.DLN = fsub(sumN, oldN); r.DLT1 = fsub(sumT1, oldT1); r.DLT2 = fsub(sumT2, oldT2);
    return r;
}

// c (1 - 1e-4) rounded down and c (1 + 2e-3) rounded up, c = sqrt(2) * (double)0.1f = 0.14142135834465...
#define NANS_KFRIC_FREE 0.1414071f
#define NANS_KFRIC_SAT 0.1417043f

// Requires: no NaN among the increments.
__device__ __forceinline__ AccumDeltas accumulate_fast(float lambdaN, float lambdaT1, float lambdaT2)
{
    AccumDeltas r;
    if (!(lambdaN > 0.0f)) {           // N_k = M_k = +0 for every k: T_k = +-0, every delta is +0
        r.DLN = r.DLT1 = r.DLT2 = 0.0f;
        return r;
    }
    const float a1 = fabsf(lambdaT1), a2 = fabsf(lambdaT2);
    const float lim_free = fmul(NANS_KFRIC_FREE, lambdaN), lim_sat = fmul(NANS_KFRIC_SAT, lambdaN);
    const bool free1 = a1 <= lim_free, free2 = a2 <= lim_free;
    const bool ok = lambdaN >= 1e-30f && lambdaN <= 1e30f && (free1 || a1 >= lim_sat) && (free2 || a2 >= lim_sat);
    if (!ok) {
#ifdef NANS_ACCUM_ON_FALLBACK
        NANS_ACCUM_ON_FALLBACK;
#endif
        return accumulate_pipelined(lambdaN, lambdaT1, lambdaT2);
    }
    // three independent chains of plain additions (a saturated tangent's chain is computed and ignored: no branch)
    float n = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll 23
    for (int it = 0; it < 69; ++it) {
        n = fadd(n, lambdaN);
        t1 = fadd(t1, lambdaT1);
        t2 = fadd(t2, lambdaT2);
    }
    const float n70 = fadd(n, lambdaN);
    r.DLN = fsub(n70, n);
    r.DLT1 = fsub(fadd(t1, lambdaT1), t1);
    r.DLT2 = fsub(fadd(t2, lambdaT2), t2);
    if (!(free1 && free2)) {
        const float m69 = friction_bound(n), m70 = friction_bound(n70);
        if (!free1) r.DLT1 = lambdaT1 < 0.0f ? fsub(-m70, -m69) : fsub(m70, m69);
        if (!free2) r.DLT2 = lambdaT2 < 0.0f ? fsub(-m70, -m69) : fsub(m70, m69);
    }
    return r;
}

}  // namespace nans
