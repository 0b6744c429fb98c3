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
// accumulate_fast      same bits without the two fp64 conversions per iteration that bound the literal loop
//                      (XU pipe; the solver's critical path is depth x this loop).  M_k is only needed exactly
//                      where it decides a result bit, so iterations 1..68 carry each tangent sum as an
//                      INTERVAL [tl, th] under the rigorous fp32 bounds
//                          lo_k = fl(cLo * N_k) <= M_k <= fl(cHi * N_k) = hi_k,
//                      cLo / cHi = the floats just below / above the fp64 friction constant (cLo * N_k is
//                      exact in fp64 and both roundings are monotone, hence the inequalities), and iterations
//                      69 and 70 use the exact M_k.  fl(x + LT) and the clamp are monotone in x, the clamp is
//                      monotone in M on either side of zero, so the true sum always lies in the interval.  A
//                      contact is in practice either never clamped (the interval stays a point) or clamped
//                      every iteration (it collapses to +-M_69, +-M_70 at the end); if an interval is still
//                      open after iteration 70, or both sums are zero (sign of zero undecided), the literal
//                      loop is run instead.  Closed forms: LN <= 0 keeps N = M = 0 and gives all deltas +0;
//                      LT = +-0 keeps T = +0.
// tests/test_solver_accum_host.py compiles this header for the host and compares the forms bit for bit
// over random, borderline and non-finite increments.
//
// Measured (B200, 1 M-cube pile, 0 fallbacks in 1.03 M contacts): the fast form is NOT faster in the kernel --
// apply 3859 vs 3499 cycles per contact.  The hand-pipelined form already hides the fp64 conversions; what
// bounds both is the FADD -> FMNMX -> FMNMX chain of a tangent sum (the same in both) and issue slots, of
// which the interval form needs 15 per iteration against 11.  It stays as the host-checked alternative.
#pragma once
#include "nans_math.cuh"

namespace nans {

#ifndef NANS_ACCUM_FAST
#define NANS_ACCUM_FAST 0   // measured on the 1 M-cube pile: fast form 0.385 ms solver stage, pipelined form 0.372 ms (see below)
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

// one tangent sum as an interval: the true sum stays inside [tl, th]
struct TanInterval {
    float tl, th;
    __device__ __forceinline__ void step(float lt, float lo, float hi)
    {
        tl = fminf(fmaxf(fadd(tl, lt), -hi), lo);
        th = fmaxf(fminf(fadd(th, lt), hi), -lo);
    }
};

// true: DLT is decided.  (old, now) = the intervals after iterations 69 and 70.
__device__ __forceinline__ bool tangent_delta(float lt, const TanInterval &old, const TanInterval &now, float &DLT)
{
    DLT = fsub(now.tl, old.tl);
    if (lt == 0.0f) { DLT = 0.0f; return true; }     // T stays +0 (NaN increments never get here)
    // a point interval away from zero has one bit pattern; x - (+-0) and (+-0) - x do not depend on the zero's sign
    return old.tl == old.th && now.tl == now.th && (old.tl != 0.0f || now.tl != 0.0f);
}

// Requires: no NaN among the increments.
__device__ __forceinline__ AccumDeltas accumulate_fast(float lambdaN, float lambdaT1, float lambdaT2)
{
    AccumDeltas r;
    if (!(lambdaN > 0.0f)) {           // N_k = M_k = +0 for every k: T_k = +-0, every delta is +0
        r.DLN = r.DLT1 = r.DLT2 = 0.0f;
        return r;
    }
    // lambdaN > 0: the sum never goes negative, the max(., 0) of the reference is the identity
    float sumN = 0.f;
    TanInterval t1 = {0.f, 0.f}, t2 = {0.f, 0.f};
#pragma unroll 4
    for (int it = 0; it < 68; ++it) {
        sumN = fadd(sumN, lambdaN);
        const float lo = fmul(NANS_KFRIC_LO, sumN), hi = fmul(NANS_KFRIC_HI, sumN);
        t1.step(lambdaT1, lo, hi);
        t2.step(lambdaT2, lo, hi);
    }
    const float n69 = fadd(sumN, lambdaN), n70 = fadd(n69, lambdaN);
    const float m69 = friction_bound(n69), m70 = friction_bound(n70);
    t1.step(lambdaT1, m69, m69); t2.step(lambdaT2, m69, m69);
    const TanInterval o1 = t1, o2 = t2;
    t1.step(lambdaT1, m70, m70); t2.step(lambdaT2, m70, m70);
    r.DLN = fsub(n70, n69);
    const bool ok1 = tangent_delta(lambdaT1, o1, t1, r.DLT1);
    const bool ok2 = tangent_delta(lambdaT2, o2, t2, r.DLT2);
    if (!(ok1 && ok2)) {
#ifdef NANS_ACCUM_ON_FALLBACK
        NANS_ACCUM_ON_FALLBACK;
#endif
        return accumulate_literal(lambdaN, lambdaT1, lambdaT2);
    }
    return r;
}

}  // namespace nans
