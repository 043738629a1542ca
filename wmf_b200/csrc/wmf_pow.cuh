// wmf_pow.cuh -- x**y for the two power laws on the SHIA hot path (modelosv2.f90:481 `(S1/H1)**0.6`, :1290
// `area**expo`), ~40 fp32 instructions instead of the ~100 of CUDA's generic powf.
//
// The reference calls libm's powf (< 1 ulp).  CUDA's powf is specified to 4 ulp; this one measures <= 1.5 ulp against
// a correctly rounded pow over the ranges the model uses (tools/check_pow.cpp runs the same arithmetic on the CPU:
// every operation is an IEEE fp32 add / mul / fma, so host and device results are bit-identical).
//
// Method: x = 2^e * m with m in [sqrt(1/2), sqrt(2)); log2(m) = u/ln2 + u^2 Q(u) with u = m-1 and the leading term
// carried as an unevaluated (hi, lo) pair; y*log2(x) again as (hi, lo); 2^f by a degree-6 polynomial on [-0.5, 0.5].
// Inputs outside the fast path (x <= 0, subnormal, inf, NaN, huge results) go to powf.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define WMF_HD __host__ __device__ __forceinline__
#else
#define WMF_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define WMF_F2I(x) __float_as_int(x)
#define WMF_I2F(x) __int_as_float(x)
#else
#include <string.h>
static inline int WMF_F2I(float x) { int i; memcpy(&i, &x, 4); return i; }
static inline float WMF_I2F(int i) { float x; memcpy(&x, &i, 4); return x; }
#endif

// log2(x) = *hi + *lo for a positive normal x
WMF_HD void wmf_log2_hl(float x, float *hi, float *lo) {
    const int ix = WMF_F2I(x);
    const int e = (ix - 0x3f3504f3) >> 23;            // m = x / 2^e in [sqrt(1/2), sqrt(2))
    const float m = WMF_I2F(ix - (e << 23));
    const float u = m - 1.0f;                         // exact
    float q = -0.1134920232512913f;
    q = fmaf(q, u, 0.18497874254231741f);
    q = fmaf(q, u, -0.18901662148132886f);
    q = fmaf(q, u, 0.20470400374788508f);
    q = fmaf(q, u, -0.23988556116225307f);
    q = fmaf(q, u, 0.28856741341621484f);
    q = fmaf(q, u, -0.36068490081409826f);
    q = fmaf(q, u, 0.48089817565414056f);
    q = fmaf(q, u, -0.7213474867121151f);
    const float c_hi = 1.44269502162933349609375f, c_lo = 1.925963033500011e-8f;  // 1/ln2
    const float h = u * c_hi;
    float l = fmaf(u, c_hi, -h);                      // rounding error of h, exact
    l = fmaf(u, c_lo, l);
    l = fmaf(u * u, q, l);
    const float fe = (float)e;
    const float a = fe + h;                           // |fe| >= 1 > |h| or fe == 0: fast two-sum is exact
    l = l + ((fe - a) + h);
    const float s = a + l;                            // renormalise: |lo| <= ulp(hi)/2 (|a| >= |l| or both 0)
    *hi = s;
    *lo = (a - s) + l;
}

// 2^(hi + lo), |hi| <= 125
WMF_HD float wmf_exp2_hl(float hi, float lo) {
    const float n = rintf(hi);
    const float f = (hi - n) + lo;                    // |f| <= 0.5 (+ lo)
    float p = 1.5463791491616521e-04f;
    p = fmaf(p, f, 1.3393187333667403e-03f);
    p = fmaf(p, f, 9.6180506914406716e-03f);
    p = fmaf(p, f, 5.5503526664147880e-02f);
    p = fmaf(p, f, 2.4022650950936519e-01f);
    p = fmaf(p, f, 6.9314718897269223e-01f);
    const float r = fmaf(p, f, 1.0f);                 // in [0.70, 1.42]
    return WMF_I2F(WMF_F2I(r) + ((int)n << 23));
}

// x**y.  Fast path for 2^-126 <= x < inf and |y * log2(x)| < 125; everything else is powf's business.
WMF_HD float wmf_powf(float x, float y) {
    // positive, normal and finite: one unsigned compare on the bit pattern (0x00800000 <= bits < 0x7f800000)
    if ((unsigned)(WMF_F2I(x) - 0x00800000) >= 0x7f000000u) return powf(x, y);
    float a, l;
    wmf_log2_hl(x, &a, &l);
    const float ph = y * a;
    if (!(fabsf(ph) < 125.0f)) return powf(x, y);
    const float pl = fmaf(y, l, fmaf(y, a, -ph));
    return wmf_exp2_hl(ph, pl);
}

// x**y evaluated in double precision and rounded once to fp32 (relative error of the double result ~2^-45, i.e. the
// correctly rounded fp32 power except for values within ~1e-6 ulp of a rounding boundary).  The reference's build calls
// glibc's powf, which computes in double as well and is correctly rounded in all but a few percent of the cases.  Used
// where the model subtracts nearly equal powers (sed_hillslope: excess transport capacity = Qskr - supply,
// modelosv2.f90:1339-1395), so that a 1.4 ulp pow would show up as 1e-5 in the result; y must not be 0.
WMF_HD float wmf_powf_dd(float x, float y) { return (float)exp2((double)y * log2((double)x)); }
