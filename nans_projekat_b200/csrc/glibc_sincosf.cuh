// glibc_sincosf.cuh — sinf/cosf bit-compatible with the host libm the reference links against.
//
// The reference's model rebuild (code/nans.cpp:1870-1881,1913-1941) goes through glm::rotate,
// which calls the HOST's sinf/cosf (glibc; imported symbols of build/nans.so).  CUDA's sinf/cosf
// round differently (and even a correctly rounded result differs from glibc's in ~1 % of inputs),
// so the vertex rebuild kernel evaluates glibc's own algorithm: the double-precision polynomial
// kernel of sysdeps/ieee754/flt-32/s_sinf.c / s_cosf.c (glibc >= 2.28, same through 2.39), in the
// FMA contraction pattern of the x86-64 `__sinf_fma` / `__cosf_fma` ifunc variants every
// FMA-capable host selects.  The constants are the published __sincosf_table / __inv_pio4.
// tests/test_host_cpu.py::test_sincos_emulation_matches_libm (tests/sincos_check.cpp compiles this header for
// the host) proves it equal to the running libm, bit for bit, over the full float range.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define NANS_HD __host__ __device__ __forceinline__
#define NANS_FMA(a, b, c) __fma_rn((a), (b), (c))
#define NANS_DMUL(a, b) __dmul_rn((a), (b))
#define NANS_D2F(a) __double2float_rn(a)
#elif defined(__CUDACC__)
#include <math.h>
#define NANS_HD __host__ __device__ __forceinline__
#define NANS_FMA(a, b, c) fma((a), (b), (c))
#define NANS_DMUL(a, b) ((a) * (b))
#define NANS_D2F(a) ((float)(a))
#else
#include <math.h>
#define NANS_HD static inline
#define NANS_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define NANS_DMUL(a, b) ((a) * (b))
#define NANS_D2F(a) ((float)(a))
#endif

namespace nans_glibc {

// __sincosf_table[q]: c0..c4 cosine, s1..s3 sine polynomial; q = 1 is the negated-cosine copy
struct SinCosTab { double c0, c1, c2, c3, c4, s1, s2, s3; };

NANS_HD SinCosTab tab(int q)
{
    SinCosTab t;
    const double sg = q ? -1.0 : 1.0;
    t.c0 = sg * 0x1p0;
    t.c1 = sg * -0x1.ffffffd0c621cp-2;
    t.c2 = sg * 0x1.55553e1068f19p-5;
    t.c3 = sg * -0x1.6c087e89a359dp-10;
    t.c4 = sg * 0x1.99343027bf8c3p-16;
    t.s1 = -0x1.555545995a603p-3;
    t.s2 = 0x1.1107605230bc4p-7;
    t.s3 = -0x1.994eb3774cf24p-13;
    return t;
}

NANS_HD float poly(double x, double x2, const SinCosTab &p, int n)
{
    if ((n & 1) == 0) {
        double s1 = NANS_FMA(x2, p.s3, p.s2);
        double x3 = NANS_DMUL(x2, x);
        double x7 = NANS_DMUL(x2, x3);
        double s = NANS_FMA(x3, p.s1, x);
        return NANS_D2F(NANS_FMA(s1, x7, s));
    } else {
        double x4 = NANS_DMUL(x2, x2);
        double c1 = NANS_FMA(x2, p.c1, p.c0);
        double c2 = NANS_FMA(x2, p.c4, p.c3);
        double x6 = NANS_DMUL(x2, x4);
        double c = NANS_FMA(x4, p.c2, c1);
        return NANS_D2F(NANS_FMA(c2, x6, c));
    }
}

NANS_HD uint32_t fbits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}

NANS_HD double sign_of_quadrant(int q) { return (q == 1 || q == 2) ? -1.0 : 1.0; }  // {1,-1,-1,1}

NANS_HD uint32_t inv_pio4(int i)
{
    const uint32_t t[24] = {0xa2u,       0xa2f9u,     0xa2f983u,   0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u,
                            0x6e4e4415u, 0x4e441529u, 0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u,
                            0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, 0xf534ddc0u, 0x34ddc0dbu, 0xddc0db62u,
                            0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, 0x993c4390u, 0x3c439041u};
    return t[i];
}

// reduce_fast: |x| < 120
NANS_HD double reduce_fast(double x, int *np)
{
    double r = NANS_DMUL(x, 0x1.45F306DC9C883p+23);
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return NANS_FMA(-(double)n, 0x1.921FB54442D18p0, x);
}

// reduce_large: 120 <= |x| < inf
NANS_HD double reduce_large(uint32_t xi, int *np)
{
    const int base = (xi >> 26) & 15;
    const int shift = (xi >> 23) & 7;
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    res0 = (uint32_t)(xi * inv_pio4(base));
    res1 = (uint64_t)xi * inv_pio4(base + 4);
    res2 = (uint64_t)xi * inv_pio4(base + 8);
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    double x = (double)(int64_t)res0;
    *np = (int)n;
    return NANS_DMUL(x, 0x1.921FB54442D18p-62);
}

NANS_HD float sinf_glibc(float y)
{
    const uint32_t bits = fbits(y);
    const uint32_t top = (bits >> 20) & 0x7ff;
    double x = (double)y;
    int n;
    if (top < 0x3f4) {                         // |y| < pi/4 (abstop12)
        if (top < 0x398) return y;             // |y| < 2^-12
        return poly(x, NANS_DMUL(x, x), tab(0), 0);
    } else if (top < 0x42f) {                  // |y| < 120
        x = reduce_fast(x, &n);
        double s = sign_of_quadrant(n & 3);
        return poly(NANS_DMUL(x, s), NANS_DMUL(x, x), tab((n & 2) ? 1 : 0), n);
    } else if (top < 0x7f8) {
        const int sign = (int)(bits >> 31);
        x = reduce_large(bits, &n);
        double s = sign_of_quadrant((n + sign) & 3);
        return poly(NANS_DMUL(x, s), NANS_DMUL(x, x), tab(((n + sign) & 2) ? 1 : 0), n);
    }
    return y - y;                              // inf/NaN -> NaN
}

NANS_HD float cosf_glibc(float y)
{
    const uint32_t bits = fbits(y);
    const uint32_t top = (bits >> 20) & 0x7ff;
    double x = (double)y;
    int n;
    if (top < 0x3f4) {
        if (top < 0x398) return 1.0f;
        return poly(x, NANS_DMUL(x, x), tab(0), 1);
    } else if (top < 0x42f) {
        x = reduce_fast(x, &n);
        double s = sign_of_quadrant(n & 3);
        return poly(NANS_DMUL(x, s), NANS_DMUL(x, x), tab((n & 2) ? 1 : 0), n ^ 1);
    } else if (top < 0x7f8) {
        const int sign = (int)(bits >> 31);
        x = reduce_large(bits, &n);
        double s = sign_of_quadrant((n + sign) & 3);
        return poly(NANS_DMUL(x, s), NANS_DMUL(x, x), tab(((n + sign) & 2) ? 1 : 0), n ^ 1);
    }
    return y - y;
}

}  // namespace nans_glibc
