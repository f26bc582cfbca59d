/* dm_math.h -- deterministic fp64 elementary functions, one source for host C and sm_100a device code.
 *
 * Why: the reference's generated problem code calls libm (sin, cos, asin ...; iLQG_func.tem:5-7 adds sec/csc).
 * glibc's libm and CUDA's libdevice do not round identically, and the car benchmark is numerically chaotic
 * (SURVEY.md 7.3-1): one ulp in a cost flips the accepted line-search step a dozen iterations later.  The
 * generated problem code on BOTH sides (reference-ABI C for the oracle, __device__ code for the kernels)
 * therefore calls these functions, which use only IEEE +,-,*,/,sqrt and integer bit operations.  Compiled with
 * `gcc -ffp-contract=off` and `nvcc -fmad=false` they return bit-identical results on host and device.
 *
 * Algorithms: the classic published fdlibm scheme (Cody-Waite three-stage pi/2 reduction, degree-13/14 minimax
 * kernels on [-pi/4, pi/4], evaluated straight-line in dm_sincos; rational approximation for asin/acos).  Accuracy < 1 ulp on the supported range
 * (checked against libm in tests/test_dm_math.py).  |x| >= 2^20*pi/2 for sin/cos returns NaN: the generated
 * NaN/Inf guards then reject the rollout exactly as the reference's guards would for a non-finite value.
 */
#ifndef DM_MATH_H
#define DM_MATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define DM_HD __host__ __device__ __forceinline__
#else
#define DM_HD static inline
#endif

DM_HD uint64_t dm_to_bits(double v)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(v);
#else
    uint64_t b;
    memcpy(&b, &v, sizeof b);
    return b;
#endif
}

DM_HD double dm_from_bits(uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double v;
    memcpy(&v, &b, sizeof v);
    return v;
#endif
}

DM_HD uint32_t dm_hi_abs(double v) { return (uint32_t)(dm_to_bits(v) >> 32) & 0x7fffffffu; }

DM_HD double dm_sqrt(double v)
{
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(v);
#else
    return sqrt(v);
#endif
}

DM_HD double dm_nan(void) { return dm_from_bits(0x7ff8000000000000ull); }

/* sin and cos of one argument, sharing the range reduction.  Straight-line code: both polynomial kernels are always
 * evaluated and the results picked by select, so lanes of a warp never diverge on the quadrant and independent calls
 * can be interleaved by the compiler (only the rare second/third reduction stage is a branch).  Every selected value
 * is computed by exactly the operations of the classic branching formulation, so results are bit-identical to it
 * (pinned by tests/golden/dm_math.npz). */
DM_HD void dm_sincos(double x, double *sn, double *cs)
{
    const double invpio2 = 6.36619772367581382433e-01;
    const double p1 = 1.57079632673412561417e+00, p1t = 6.07710050650619224932e-11;
    const double p2 = 6.07710050630396597660e-11, p2t = 2.02226624879595063154e-21;
    const double p3 = 2.02226624871116645580e-21, p3t = 8.47842766036889956997e-32;
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const uint32_t ix = dm_hi_abs(x);
    const int small = (ix <= 0x3fe921fbu);      /* |x| <= ~pi/4: no reduction */
    const int bad = (ix >= 0x413921fbu);        /* |x| >= 2^20*pi/2, inf or nan */
    const int neg = ((int64_t)dm_to_bits(x) < 0);
    const double ax = dm_from_bits(dm_to_bits(x) & 0x7fffffffffffffffull);
    const int n = (small | bad) ? 0 : (int)(ax * invpio2 + 0.5);
    const double fn = (double)n;
    double r = ax - fn * p1;
    double w = fn * p1t;
    const int j = (int)(ix >> 20);
    double h = r - w;
    int i = j - (int)((dm_hi_abs(h) >> 20) & 0x7ffu);
    if (i > 16 && !bad) { /* second stage (rare: heavy cancellation) */
        double t = r;
        w = fn * p2;
        r = t - w;
        w = fn * p2t - ((t - r) - w);
        h = r - w;
        i = j - (int)((dm_hi_abs(h) >> 20) & 0x7ffu);
        if (i > 49) { /* third stage */
            t = r;
            w = fn * p3;
            r = t - w;
            w = fn * p3t - ((t - r) - w);
            h = r - w;
        }
    }
    const double l = (r - h) - w;
    const double y0 = small ? x : (neg ? -h : h);
    const double y1 = small ? 0.0 : (neg ? -l : l);
    const int q = (neg ? -n : n) & 3;
    /* sine kernel, both published variants (with / without tail), selected like the branching code does */
    const double z = y0 * y0;
    const double v = z * y0;
    const double rs = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    const double ks_notail = y0 + v * (S1 + z * rs);
    const double ks_tail = y0 - ((z * (0.5 * y1 - v * rs) - y1) - v * S1);
    const double ks = small ? ks_notail : ks_tail;
    /* cosine kernel */
    const uint32_t iy = dm_hi_abs(y0);
    const double rc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    const double xy = z * rc - y0 * y1;
    const double kc_small = 1.0 - (0.5 * z - xy);
    const double qx = (iy > 0x3fe90000u) ? 0.28125 : dm_from_bits((uint64_t)(uint32_t)(iy - 0x00200000u) << 32);
    const double hz = 0.5 * z - qx;
    const double kc_big = (1.0 - qx) - (hz - xy);
    const double kc = (iy < 0x3fd33333u) ? kc_small : kc_big;
    /* quadrant */
    const double s_sel = (q & 1) ? kc : ks;
    const double c_sel = (q & 1) ? ks : kc;
    const double s_val = (q & 2) ? -s_sel : s_sel;
    const double c_val = ((q + 1) & 2) ? -c_sel : c_sel;
    *sn = bad ? dm_nan() : s_val;
    *cs = bad ? dm_nan() : c_val;
}

DM_HD double dm_sin(double x)
{
    double s, c;
    dm_sincos(x, &s, &c);
    return s;
}

DM_HD double dm_cos(double x)
{
    double s, c;
    dm_sincos(x, &s, &c);
    return c;
}

DM_HD double dm_tan(double x) { return dm_sin(x) / dm_cos(x); }

DM_HD double dm_asin(double x)
{
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17,
                 pio4_hi = 7.85398163397448278999e-01;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05;
    const double qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    uint32_t ix = dm_hi_abs(x);
    int neg = ((int64_t)dm_to_bits(x) < 0);
    double t, w, p, q;
    if (ix >= 0x3ff00000u) { /* |x| >= 1 */
        if (x == 1.0 || x == -1.0)
            return x * pio2_hi + x * pio2_lo;
        return dm_nan();
    }
    if (ix < 0x3fe00000u) { /* |x| < 0.5 */
        if (ix < 0x3e400000u)
            return x;
        t = x * x;
        p = t * (pS0 + t * (pS1 + t * (pS2 + t * (pS3 + t * (pS4 + t * pS5)))));
        q = 1.0 + t * (qS1 + t * (qS2 + t * (qS3 + t * qS4)));
        w = p / q;
        return x + x * w;
    }
    double ax = dm_from_bits(dm_to_bits(x) & 0x7fffffffffffffffull);
    w = 1.0 - ax;
    t = w * 0.5;
    p = t * (pS0 + t * (pS1 + t * (pS2 + t * (pS3 + t * (pS4 + t * pS5)))));
    q = 1.0 + t * (qS1 + t * (qS2 + t * (qS3 + t * qS4)));
    double s = dm_sqrt(t);
    if (ix >= 0x3fef3333u) { /* |x| > 0.975 */
        w = p / q;
        t = pio2_hi - (2.0 * (s + s * w) - pio2_lo);
    } else {
        w = dm_from_bits(dm_to_bits(s) & 0xffffffff00000000ull);
        double c = (t - w * w) / (s + w);
        double r = p / q;
        p = 2.0 * s * r - (pio2_lo - 2.0 * c);
        q = pio4_hi - 2.0 * w;
        t = pio4_hi - (p - q);
    }
    return neg ? -t : t;
}

DM_HD double dm_acos(double x)
{
    const double pi = 3.14159265358979311600e+00, pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05;
    const double qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    uint32_t ix = dm_hi_abs(x);
    int neg = ((int64_t)dm_to_bits(x) < 0);
    double z, p, q, r, s, w;
    if (ix >= 0x3ff00000u) {
        if (x == 1.0) return 0.0;
        if (x == -1.0) return pi + 2.0 * pio2_lo;
        return dm_nan();
    }
    if (ix < 0x3fe00000u) { /* |x| < 0.5 */
        if (ix <= 0x3c600000u) return pio2_hi + pio2_lo;
        z = x * x;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if (neg) { /* x < -0.5 */
        z = (1.0 + x) * 0.5;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        s = dm_sqrt(z);
        r = p / q;
        w = r * s - pio2_lo;
        return pi - 2.0 * (s + w);
    }
    z = (1.0 - x) * 0.5; /* x > 0.5 */
    s = dm_sqrt(z);
    double df = dm_from_bits(dm_to_bits(s) & 0xffffffff00000000ull);
    double c = (z - df * df) / (s + df);
    p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    r = p / q;
    w = r * s + c;
    return 2.0 * (df + w);
}

DM_HD double dm_fabs(double v) { return dm_from_bits(dm_to_bits(v) & 0x7fffffffffffffffull); }

DM_HD int dm_isfinite(double v) { return ((dm_to_bits(v) >> 52) & 0x7ffu) != 0x7ffu; }

#endif /* DM_MATH_H */
