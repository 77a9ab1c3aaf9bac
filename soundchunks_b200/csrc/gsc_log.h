/*
 * gsc_log.h -- natural logarithm of a double, correctly rounded (to within 2^-100 of the exact value before the
 * final rounding), written in plain IEEE-754 double operations plus fma(), so that the CUDA kernels and any C host
 * code that includes this header produce the SAME double bit for bit.
 *
 * Why: the cepstral half of the chunk features goes through log10 (enc:316-318, math.log10 = ln(x) * const).  The
 * reference runs FreePascal's ln (x87 fyl2x, extended precision, rounded to Double on the store: the correctly
 * rounded Double except in ~2^-10 of the cases); CUDA's log() and glibc's log() are both "< 1 ulp" functions that
 * differ from each other in the last bit now and then, which made feature bits depend on the platform.  A shared
 * correctly-rounded routine removes the platform from the result and is the closest portable stand-in for the
 * x87 value.  tests/test_log_cr.py checks it against 400-bit mpmath.
 *
 * Method: x = 2^e * m with m in [sqrt(1/2), sqrt(2));  log m = 2 atanh(s), s = (m-1)/(m+1) (|s| <= 0.1716), the
 * odd series sum s^(2k+1)/(2k+1) evaluated by Horner in z = s^2 -- terms k >= 11 in double, k = 10..0 in
 * double-double -- and e*ln2 added in double-double.  No table, no branch on the data apart from the range split.
 *
 * The translation unit must not contract a*b+c (nvcc -fmad=false, gcc -ffp-contract=off); fma() is explicit.
 */
#ifndef GSC_LOG_H
#define GSC_LOG_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define GSC_LOG_FN __host__ __device__ __forceinline__
#else
#define GSC_LOG_FN static inline
#endif

GSC_LOG_FN void gsc_dd_two_sum(double a, double b, double *s, double *e) {
    const double t = a + b;
    const double bb = t - a;
    *e = (a - (t - bb)) + (b - bb);
    *s = t;
}
GSC_LOG_FN void gsc_dd_fast_two_sum(double a, double b, double *s, double *e) {   /* |a| >= |b| or a == 0 */
    const double t = a + b;
    *e = b - (t - a);
    *s = t;
}
GSC_LOG_FN void gsc_dd_mul(double ah, double al, double bh, double bl, double *ph, double *pl) {
    const double p = ah * bh;
    double e = fma(ah, bh, -p);
    const double c1 = ah * bl, c2 = al * bh;
    e = e + (c1 + c2);
    gsc_dd_fast_two_sum(p, e, ph, pl);
}
GSC_LOG_FN void gsc_dd_add(double ah, double al, double bh, double bl, double *sh, double *sl) {
    double s, e;
    gsc_dd_two_sum(ah, bh, &s, &e);
    e = e + (al + bl);
    gsc_dd_fast_two_sum(s, e, sh, sl);
}

GSC_LOG_FN double gsc_log_cr(double x) {
    /* 1/(2k+1) as double-double, k = 0..20 (generated with mpmath at 400 bits) */
    const double ch[21] = {
        0x1.0000000000000p+0, 0x1.5555555555555p-2, 0x1.999999999999ap-3, 0x1.2492492492492p-3,
        0x1.c71c71c71c71cp-4, 0x1.745d1745d1746p-4, 0x1.3b13b13b13b14p-4, 0x1.1111111111111p-4,
        0x1.e1e1e1e1e1e1ep-5, 0x1.af286bca1af28p-5, 0x1.8618618618618p-5, 0x1.642c8590b2164p-5,
        0x1.47ae147ae147bp-5, 0x1.2f684bda12f68p-5, 0x1.1a7b9611a7b96p-5, 0x1.0842108421084p-5,
        0x1.f07c1f07c1f08p-6, 0x1.d41d41d41d41dp-6, 0x1.bacf914c1bad0p-6, 0x1.a41a41a41a41ap-6,
        0x1.8f9c18f9c18fap-6};
    const double cl[11] = {
        0x0.0p+0, 0x1.5555555555555p-56, -0x1.999999999999ap-57, 0x1.2492492492492p-57,
        0x1.c71c71c71c71cp-58, -0x1.745d1745d1746p-59, -0x1.3b13b13b13b14p-58, 0x1.1111111111111p-60,
        0x1.e1e1e1e1e1e1ep-61, 0x1.af286bca1af28p-59, 0x1.8618618618618p-59};
    const double ln2h = 0x1.62e42fefa39efp-1, ln2l = 0x1.abc9e3b39803fp-56;

    uint64_t b;
    memcpy(&b, &x, 8);
    if ((b >> 63) != 0 || (b >> 52) == 0x7ff || (b << 1) == 0) {   /* negative, zero, inf, NaN */
        if ((b << 1) == 0) return -INFINITY;
        if (x != x) return x;
        if ((b >> 63) != 0) return NAN;
        return x;                                                     /* +inf */
    }
    int e = 0;
    if ((b >> 52) == 0) {            /* subnormal: scale by 2^54 (exact) */
        x = x * 0x1p54;
        memcpy(&b, &x, 8);
        e = -54;
    }
    e += (int)(b >> 52) - 1023;
    uint64_t mb = b & 0x000fffffffffffffull;
    if (mb > 0x6a09e667f3bcdull) { e += 1; mb |= 0x3fe0000000000000ull; }   /* m in [sqrt2/2 .. 1) */
    else mb |= 0x3ff0000000000000ull;                                        /* m in [1 .. sqrt2]   */
    double m;
    memcpy(&m, &mb, 8);

    /* s = (m - 1) / (m + 1) in double-double: m - 1 is exact, m + 1 exact as a pair */
    const double a = m - 1.0;
    double bh, bl;
    gsc_dd_two_sum(m, 1.0, &bh, &bl);
    const double q1 = a / bh;
    double r = fma(-q1, bh, a);
    r = r - q1 * bl;
    const double q2 = r / bh;
    double sh, sl;
    gsc_dd_fast_two_sum(q1, q2, &sh, &sl);

    double zh, zl;
    gsc_dd_mul(sh, sl, sh, sl, &zh, &zl);
    /* P(z) = sum_k z^k / (2k+1): tail in double, head in double-double */
    double p = ch[20];
    for (int k = 19; k >= 11; --k) p = ch[k] + zh * p;
    double ph = p, pl = 0.0;
    for (int k = 10; k >= 0; --k) {
        double th, tl;
        gsc_dd_mul(zh, zl, ph, pl, &th, &tl);
        gsc_dd_add(ch[k], cl[k], th, tl, &ph, &pl);
    }
    double lh, ll;
    gsc_dd_mul(sh, sl, ph, pl, &lh, &ll);
    lh = lh * 2.0; ll = ll * 2.0;                                   /* log m */

    const double ed = (double)e;
    double eh = ed * ln2h;
    double el = fma(ed, ln2h, -eh);
    el = el + ed * ln2l;
    double yh, yl;
    gsc_dd_add(eh, el, lh, ll, &yh, &yl);
    return yh;
}

#endif /* GSC_LOG_H */
