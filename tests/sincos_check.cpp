// Host check: nans_glibc::sinf_glibc/cosf_glibc (the code the vertex-rebuild kernel runs) against
// the running libm's sinf/cosf, bit for bit.  argv[1] = number of random samples (plus a dense
// sweep of every 2^k-th float).  Prints "mismatch_sin mismatch_cos total".
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../nans_projekat_b200/csrc/glibc_sincosf.cuh"

static inline uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float fromb(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main(int argc, char **argv)
{
    long n = argc > 1 ? atol(argv[1]) : 1000000;
    unsigned stride = argc > 2 ? (unsigned)atol(argv[2]) : 4099u;
    long bad_s = 0, bad_c = 0, total = 0;
    volatile float (*psin)(float) = (volatile float (*)(float))sinf; (void)psin;
    // dense sweep over all finite floats with a prime stride (both signs)
    for (uint64_t u = 0; u < 0x7f800000ull; u += stride) {
        for (int sg = 0; sg < 2; ++sg) {
            float x = fromb((uint32_t)u | (sg ? 0x80000000u : 0u));
            float a = sinf(x), b = nans_glibc::sinf_glibc(x);
            float c = cosf(x), d = nans_glibc::cosf_glibc(x);
            bad_s += bits(a) != bits(b);
            bad_c += bits(c) != bits(d);
            if ((bits(a) != bits(b) || bits(c) != bits(d)) && bad_s + bad_c < 6)
                fprintf(stderr, "x=%a sinf=%a emu=%a cosf=%a emu=%a\n", x, a, b, c, d);
            ++total;
        }
    }
    // random samples in the ranges the step actually visits (radians(Angles))
    uint64_t s = 88172645463325252ull;
    for (long i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        float r = (float)((s >> 11) * (1.0 / 9007199254740992.0)) * 2.0f - 1.0f;
        float scale = (i % 4 == 0) ? 0.05f : (i % 4 == 1) ? 1.0f : (i % 4 == 2) ? 10.0f : 200.0f;
        float x = r * scale;
        bad_s += bits(sinf(x)) != bits(nans_glibc::sinf_glibc(x));
        bad_c += bits(cosf(x)) != bits(nans_glibc::cosf_glibc(x));
        ++total;
    }
    printf("%ld %ld %ld\n", bad_s, bad_c, total);
    return 0;
}
