// np_host_shim.cpp — TEST INFRASTRUCTURE.  Compiles the DEVICE narrowphase (csrc/narrowphase.cuh, the code
// the CUDA kernels run) as host C++ so its algorithm can be checked against the oracle without a GPU:
// one "thread" (threadIdx.x = 0, kNpThreads = 1), __shared__ = static storage, the __f*_rn intrinsics =
// plain IEEE fp32 operations (build with -ffp-contract=off, SSE2: never fused, no x87).
// Built by tests/test_np_host.py into tests/_build/; never loaded by the product path.
#include <cuda_runtime.h>   // host side: float4, make_float4, uint2 ...
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
using std::max;
using std::min;

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
#undef __shared__
#define __shared__ static
#undef __noinline__
#define __noinline__
static const struct { unsigned x, y, z; } threadIdx = {0, 0, 0};
#define NANS_NP_THREADS 1
#define NANS_HOST_SHIM 1

#include "../nans_projekat_b200/csrc/narrowphase.cuh"

using namespace nans;

static void load_box_host(int side, const float *v24)
{
    float v[24];
    for (int q = 0; q < 24; ++q) v[q] = v24[q];
    NpShapes::store_box(side, v);
}

template <bool AS, bool BS>
static NpResult run(NpShapes &S, int &ovf, int &mf)
{
    static EpaArena E;
    return check_collision<AS, BS>(S, E, ovf, mf);
}

// same argument meaning as nans_check_collision_batch (include/nans_b200.h); returns the overflow bits
extern "C" int np_host_check_collision_batch(int n, const int32_t *type, const float *pos_a, const float *verts_a,
                                             const float *rad_a, const float *pos_b, const float *verts_b,
                                             const float *rad_b, int32_t *hit, int32_t *gjk, float *N, float *PA,
                                             float *PB, int32_t *max_faces)
{
    int ovf = 0, mf = 0;
    for (int p = 0; p < n; ++p) {
        const int t = type[p];
        const bool as = (t == NANS_SS || t == NANS_SF), bs = (t == NANS_CS || t == NANS_SS);
        NpShapes S;
        S.posA = V3(pos_a[3 * p], pos_a[3 * p + 1], pos_a[3 * p + 2]); S.radA = rad_a[p];
        S.posB = V3(pos_b[3 * p], pos_b[3 * p + 1], pos_b[3 * p + 2]); S.radB = rad_b[p];
        if (!as) load_box_host(0, verts_a + 24 * (size_t)p);
        if (!bs) load_box_host(1, verts_b + 24 * (size_t)p);
        NpResult r;
        if (!as && !bs) r = run<false, false>(S, ovf, mf);
        else if (!as && bs) r = run<false, true>(S, ovf, mf);
        else if (as && !bs) r = run<true, false>(S, ovf, mf);
        else r = run<true, true>(S, ovf, mf);
        hit[p] = r.hit; gjk[p] = r.gjk;
        N[3 * p] = r.N.x; N[3 * p + 1] = r.N.y; N[3 * p + 2] = r.N.z;
        PA[3 * p] = r.PA.x; PA[3 * p + 1] = r.PA.y; PA[3 * p + 2] = r.PA.z;
        PB[3 * p] = r.PB.x; PB[3 * p + 1] = r.PB.y; PB[3 * p + 2] = r.PB.z;
    }
    if (max_faces) *max_faces = mf;
    return ovf;
}

// the vertex hash that gates EPA's class scan: equal vectors (+0 == -0 included) must have equal hashes
extern "C" unsigned np_host_hash16(float x, float y, float z) { return epa_hash16(V3(x, y, z)); }
