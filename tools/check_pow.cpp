// check_pow.cpp -- CPU check of wmf_b200/csrc/wmf_pow.cuh (the same IEEE fp32 arithmetic the GPU executes) and of the
// division-by-constant sequences used in shia.cu.  Build & run:  g++ -O2 -ffp-contract=off -o /tmp/check_pow tools/check_pow.cpp && /tmp/check_pow
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cmath>
#include <random>

#include "../wmf_b200/csrc/wmf_pow.cuh"

static double ulp_err(float got, double want) {
    const float w = (float)want;
    const double u = std::fabs((double)std::nextafterf(w, INFINITY) - (double)w);
    return std::fabs((double)got - want) / u;
}

// x / c for a constant c: q0 = x * (1/c); r = fma(-c, q0, x); q = fma(r, 1/c, q0)
static inline float div_const(float x, float c, float rc) {
    const float q0 = x * rc;
    const float r = fmaf(-c, q0, x);
    return fmaf(r, rc, q0);
}

int main() {
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double worst06 = 0, worstg = 0, worst_libm = 0;
    // (S1/H1)**0.6 : x in (0, 1]
    for (long i = 0; i < 20000000; ++i) {
        const float x = (float)std::exp(-20.0 * U(rng) * U(rng));
        const double w = std::pow((double)x, (double)0.6f);
        worst06 = std::fmax(worst06, ulp_err(wmf_powf(x, 0.6f), w));
        if (i < 4000000) worst_libm = std::fmax(worst_libm, ulp_err(powf(x, 0.6f), w));
    }
    // area**expo : x in [1e-12, 1e6], y in [0.05, 3]
    for (long i = 0; i < 20000000; ++i) {
        const float x = (float)std::exp(std::log(1e-12) + U(rng) * (std::log(1e6) - std::log(1e-12)));
        const float y = (float)(0.05 + 2.95 * U(rng));
        const double w = std::pow((double)x, (double)y);
        if (w < 1e-37 || w > 1e37) continue;
        worstg = std::fmax(worstg, ulp_err(wmf_powf(x, y), w));
    }
    printf("wmf_powf(x,0.6) x in (0,1]: worst %.3f ulp   (glibc powf on the same points: %.3f ulp)\n", worst06, worst_libm);
    printf("wmf_powf(x,y) x in [1e-12,1e6], y in [0.05,3]: worst %.3f ulp\n", worstg);
    printf("specials: pow(1,0.6)=%g pow(0,0.6)=%g pow(-1,0.6)=%g pow(inf,0.6)=%g pow(1e-45,0.6)=%g (powf %g)\n", wmf_powf(1.f, 0.6f),
           wmf_powf(0.f, 0.6f), wmf_powf(-1.f, 0.6f), wmf_powf(INFINITY, 0.6f), wmf_powf(1e-45f, 0.6f), powf(1e-45f, 0.6f));
    // exhaustive: x/3 and x/1000 over every finite float
    long bad3 = 0, bad1000 = 0;
    const float r3 = 1.0f / 3.0f, r1000 = 1.0f / 1000.0f;
    for (uint64_t b = 0; b < (1ull << 32); ++b) {  // ~25 s
        const float x = WMF_I2F((int)(uint32_t)b);
        if (!std::isfinite(x)) continue;
        const float a = x / 3.0f, c = x / 1000.0f;
        // the sequences are used for results in the normal range only (callers guard); skip tiny / huge quotients
        if (std::fabs(a) >= 1e-30f && std::fabs(x) < 1e37f && div_const(x, 3.0f, r3) != a) ++bad3;
        if (std::fabs(c) >= 1e-30f && std::fabs(x) < 1e37f && div_const(x, 1000.0f, r1000) != c) ++bad1000;
    }
    printf("x/3 via fma sequence: %ld mismatches; x/1000: %ld mismatches (all finite x with |q| >= 1e-30)\n", bad3, bad1000);
    return (worst06 < 2.0 && worstg < 2.0 && bad3 == 0 && bad1000 == 0) ? 0 : 1;
}
