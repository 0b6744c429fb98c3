// solver_accum_host_shim.cpp — TEST INFRASTRUCTURE.  Compiles csrc/solver_accum.cuh (the device code of the
// 70-iteration accumulation) as host C++ and exposes both forms, so the fast form can be compared with the
// literal one bit for bit over hundreds of millions of increments without a GPU.  fminf/fmaxf follow CUDA's
// semantics (a NaN argument loses, -0 < +0); fp64 multiply and the conversions are IEEE on both sides.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline float cuda_fmaxf(float a, float b)
{
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.0f && b == 0.0f) return signbit(a) ? b : a;      // +0 beats -0
    return a > b ? a : b;
}
static inline float cuda_fminf(float a, float b)
{
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.0f && b == 0.0f) return signbit(a) ? a : b;      // -0 beats +0
    return a < b ? a : b;
}
#define fmaxf cuda_fmaxf
#define fminf cuda_fminf

static long g_fallbacks = 0;
#define NANS_ACCUM_ON_FALLBACK (++g_fallbacks)
#include "../nans_projekat_b200/csrc/solver_accum.cuh"

extern "C" void accum_literal(long n, const float *ln, const float *lt1, const float *lt2, float *out)
{
    for (long i = 0; i < n; ++i) {
        const nans::AccumDeltas r = nans::accumulate_literal(ln[i], lt1[i], lt2[i]);
        out[3 * i] = r.DLN; out[3 * i + 1] = r.DLT1; out[3 * i + 2] = r.DLT2;
    }
}
// the dispatch of constraint_apply: NaN increments take the literal form.  returns the number of fallbacks
extern "C" long accum_fast(long n, const float *ln, const float *lt1, const float *lt2, float *out)
{
    g_fallbacks = 0;
    for (long i = 0; i < n; ++i) {
        const bool nan = ln[i] != ln[i] || lt1[i] != lt1[i] || lt2[i] != lt2[i];
        const nans::AccumDeltas r = nan ? nans::accumulate_literal(ln[i], lt1[i], lt2[i])
                                        : nans::accumulate_fast(ln[i], lt1[i], lt2[i]);
        out[3 * i] = r.DLN; out[3 * i + 1] = r.DLT1; out[3 * i + 2] = r.DLT2;
    }
    return g_fallbacks;
}
extern "C" void accum_pipelined(long n, const float *ln, const float *lt1, const float *lt2, float *out)
{
    for (long i = 0; i < n; ++i) {
        const bool nan = ln[i] != ln[i] || lt1[i] != lt1[i] || lt2[i] != lt2[i];
        const nans::AccumDeltas r = nan ? nans::accumulate_literal(ln[i], lt1[i], lt2[i])
                                        : nans::accumulate_pipelined(ln[i], lt1[i], lt2[i]);
        out[3 * i] = r.DLN; out[3 * i + 1] = r.DLT1; out[3 * i + 2] = r.DLT2;
    }
}
