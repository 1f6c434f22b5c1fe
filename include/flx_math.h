/*
 * flx_math.h -- pinned transcendental functions shared by host and device.
 *
 * The reference's kernels call OpenCL built-ins (sin, cos, tan, atan2, acos, pow,
 * native_sin/cos/recip) whose results are implementation-defined under
 * -cl-fast-relaxed-math (reference: src/clcontext.cpp:143-153).  To make "same seed ->
 * same path" a checkable property, every consumer in this repo -- the CUDA kernels, the
 * C restatement in oracle/, and the OpenCL-C shim that compiles the reference's own
 * kernel sources for the host (oracle/ref_shim) -- evaluates them with the functions in
 * this header.  They use only IEEE-754 double +,-,*,/,sqrt,rint and integer bit moves,
 * so gcc (-ffp-contract=off) and nvcc (-fmad=false) produce bit-identical floats.
 * Accuracy: <= 1 ulp of the correctly rounded float result (tests/test_math.py checks
 * against numpy in float64).
 *
 * C99 / C++ / CUDA compatible.  No dependency on libm except sqrt() and rint().
 */
#ifndef FLX_MATH_H
#define FLX_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FLX_HD __host__ __device__ __forceinline__
#else
#define FLX_HD static inline
#endif

#define FLX_PI_F 3.14159274101257f /* OpenCL M_PI_F; equals (float)PI of geom.h:19 */
#define FLX_2PI_F 6.2831853071795864f /* geom.h:21 M_2PI_F */
#define FLX_INV_PI_F 0.3183098861837907f /* geom.h:20 M_INV_PI */

FLX_HD int64_t flx__d2bits(double d)
{
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(d);
#else
    int64_t b;
    memcpy(&b, &d, sizeof b);
    return b;
#endif
}

FLX_HD double flx__bits2d(int64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double d;
    memcpy(&d, &b, sizeof d);
    return d;
#endif
}

/* sin and cos of r, |r| <= pi/4 (+ slack), Taylor in double; truncation < 1e-12 */
FLX_HD double flx__sin_k(double r)
{
    const double z = r * r;
    double p = 1.6059043836821613e-10;          /*  1/13! */
    p = p * z + -2.505210838544172e-08;          /* -1/11! */
    p = p * z + 2.7557319223985893e-06;          /*  1/9!  */
    p = p * z + -0.0001984126984126984;          /* -1/7!  */
    p = p * z + 0.008333333333333333;            /*  1/5!  */
    p = p * z + -0.16666666666666666;            /* -1/3!  */
    return r + r * (z * p);
}

FLX_HD double flx__cos_k(double r)
{
    const double z = r * r;
    double p = -1.1470745597729725e-11;          /* -1/14! */
    p = p * z + 2.08767569878681e-09;            /*  1/12! */
    p = p * z + -2.755731922398589e-07;          /* -1/10! */
    p = p * z + 2.48015873015873e-05;            /*  1/8!  */
    p = p * z + -0.001388888888888889;           /* -1/6!  */
    p = p * z + 0.041666666666666664;            /*  1/4!  */
    p = p * z + -0.5;                            /* -1/2!  */
    return 1.0 + z * p;
}

/* r = x - k*pi/2 with k = rint(x*2/pi); returns k mod 4 in *q. Valid for |x| < 2^30. */
FLX_HD double flx__rem_pio2(float x, int *q)
{
    const double xd = (double)x;
    const double kd = rint(xd * 0.6366197723675814);     /* 2/pi */
    const double r = (xd - kd * 1.5707963267341256) - kd * 6.077100506506192e-11; /* pi/2 = hi(33 bits) + lo */
    *q = (int)((int64_t)kd & 3);
    return r;
}

FLX_HD float flx_sinf(float x)
{
    int q;
    const double r = flx__rem_pio2(x, &q);
    double v;
    switch (q)
    {
    case 0: v = flx__sin_k(r); break;
    case 1: v = flx__cos_k(r); break;
    case 2: v = -flx__sin_k(r); break;
    default: v = -flx__cos_k(r); break;
    }
    return (float)v;
}

FLX_HD float flx_cosf(float x)
{
    int q;
    const double r = flx__rem_pio2(x, &q);
    double v;
    switch (q)
    {
    case 0: v = flx__cos_k(r); break;
    case 1: v = -flx__sin_k(r); break;
    case 2: v = -flx__cos_k(r); break;
    default: v = flx__sin_k(r); break;
    }
    return (float)v;
}

FLX_HD float flx_tanf(float x)
{
    int q;
    const double r = flx__rem_pio2(x, &q);
    const double s = flx__sin_k(r), c = flx__cos_k(r);
    return (float)((q & 1) ? (-c / s) : (s / c));
}

/* atan(a) for a in [0, 1], double.  One reduction through (a-1)/(a+1), then a degree-10
 * polynomial in a^2 fitted on [0, tan(pi/8)^2] (max error 4.5e-16). */
FLX_HD double flx__atan01(double a)
{
    double base = 0.0;
    if (a > 0.41421356237309503)
    {
        a = (a - 1.0) / (a + 1.0);
        base = 0.7853981633974483;
    }
    const double z = a * a;
    double p = 0.021396707182827036;
    p = p * z + -0.04369057478192942;
    p = p * z + 0.05695523817308785;
    p = p * z + -0.06641592711262516;
    p = p * z + 0.07690109667068769;
    p = p * z + -0.09090784287188458;
    p = p * z + 0.11111106680234124;
    p = p * z + -0.1428571419433257;
    p = p * z + 0.19999999999054235;
    p = p * z + -0.33333333333329823;
    return base + (a + a * (z * p));
}

FLX_HD double flx__atan2d(double y, double x)
{
    const double ay = y < 0.0 ? -y : y;
    const double ax = x < 0.0 ? -x : x;
    double r;
    if (ax == 0.0 && ay == 0.0)
        r = 0.0;
    else if (ay <= ax)
        r = flx__atan01(ay / ax);
    else
        r = 1.5707963267948966 - flx__atan01(ax / ay);
    if (x < 0.0 || (x == 0.0 && flx__d2bits(x) < 0))
        r = 3.141592653589793 - r;
    return (y < 0.0 || (y == 0.0 && flx__d2bits(y) < 0)) ? -r : r;
}

FLX_HD float flx_atan2f(float y, float x)
{
    if (y != y || x != x)
        return y + x;
    return (float)flx__atan2d((double)y, (double)x);
}

FLX_HD float flx_acosf(float x)
{
    const double xd = (double)x;
    const double s = (1.0 - xd) * (1.0 + xd); /* NaN for |x| > 1 through sqrt of a negative */
    return (float)flx__atan2d(sqrt(s), xd);
}

/* log2 of a positive, finite, normal double */
FLX_HD double flx__log2d(double x)
{
    int64_t b = flx__d2bits(x);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    b = (b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL;
    double m = flx__bits2d(b); /* [1,2) */
    if (m > 1.4142135623730951)
    {
        m = m * 0.5;
        e += 1;
    }
    const double s = (m - 1.0) / (m + 1.0);
    const double z = s * s;
    double p = 0.05263157894736842;     /* 1/19 */
    p = p * z + 0.058823529411764705;   /* 1/17 */
    p = p * z + 0.06666666666666667;    /* 1/15 */
    p = p * z + 0.07692307692307693;    /* 1/13 */
    p = p * z + 0.09090909090909091;    /* 1/11 */
    p = p * z + 0.1111111111111111;     /* 1/9  */
    p = p * z + 0.14285714285714285;    /* 1/7  */
    p = p * z + 0.2;                    /* 1/5  */
    p = p * z + 0.3333333333333333;     /* 1/3  */
    const double ln_m = 2.0 * (s + s * (z * p));
    return (double)e + ln_m * 1.4426950408889634; /* 1/ln 2 */
}

/* 2^t for |t| < 1000 */
FLX_HD double flx__exp2d(double t)
{
    const double kd = rint(t);
    const double f = (t - kd) * 0.6931471805599453; /* |f| <= 0.3466 */
    double p = 2.505210838544172e-08;   /* 1/11! */
    p = p * f + 2.755731922398589e-07;  /* 1/10! */
    p = p * f + 2.7557319223985893e-06; /* 1/9!  */
    p = p * f + 2.48015873015873e-05;   /* 1/8!  */
    p = p * f + 0.0001984126984126984;  /* 1/7!  */
    p = p * f + 0.001388888888888889;   /* 1/6!  */
    p = p * f + 0.008333333333333333;   /* 1/5!  */
    p = p * f + 0.041666666666666664;   /* 1/4!  */
    p = p * f + 0.16666666666666666;    /* 1/3!  */
    p = p * f + 0.5;
    p = p * f + 1.0;
    p = p * f + 1.0;
    const int64_t k = (int64_t)kd;
    return p * flx__bits2d((k + 1023) << 52);
}

/* pow(x, y) for the uses in the path: x >= 0 (albedo), y > 0 (2.2, utils.cl:139). */
FLX_HD float flx_powf(float x, float y)
{
    if (x != x || y != y)
        return x + y;
    if (x < 0.0f)
        return (float)flx__bits2d(0x7ff8000000000000LL); /* NaN */
    if (x == 0.0f)
        return (y > 0.0f) ? 0.0f : ((y == 0.0f) ? 1.0f : 1.0f / 0.0f);
    if (x > 3.4028234663852886e38f)
        return (y > 0.0f) ? x : ((y == 0.0f) ? 1.0f : 0.0f);
    double t = (double)y * flx__log2d((double)x);
    if (t > 300.0)
        t = 300.0; /* overflows to +inf when rounded to float */
    if (t < -300.0)
        t = -300.0; /* underflows to 0 */
    return (float)flx__exp2d(t);
}

/*
 * OpenCL read_imagef with CLK_NORMALIZED_COORDS_TRUE | CLK_ADDRESS_CLAMP_TO_EDGE |
 * CLK_FILTER_LINEAR on an RGBA float image (reference: src/env_map.cl:10,42), following the
 * filtering equations of the OpenCL 1.2 specification section 8.2: u = s*w, i0 = floor(u-0.5),
 * a = frac(u-0.5), texels clamped to the edge, T = (1-a)(1-b)T00 + a(1-b)T10 + (1-a)bT01 + abT11.
 * Hardware texture units use 8-bit weights; this is the fp32 form, and the evaluation order
 * below is the pinned one.
 */
FLX_HD void flx_bilinear_rgba(const float *img, int w, int h, float s, float t, float out[4])
{
    const float u = s * (float)w - 0.5f;
    const float v = t * (float)h - 0.5f;
    const float fu = floorf(u), fv = floorf(v);
    const float a = u - fu, b = v - fv;
    int i0 = (int)fu, j0 = (int)fv;
    int i1 = i0 + 1, j1 = j0 + 1;
    i0 = i0 < 0 ? 0 : (i0 > w - 1 ? w - 1 : i0);
    i1 = i1 < 0 ? 0 : (i1 > w - 1 ? w - 1 : i1);
    j0 = j0 < 0 ? 0 : (j0 > h - 1 ? h - 1 : j0);
    j1 = j1 < 0 ? 0 : (j1 > h - 1 ? h - 1 : j1);
    const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    const float *t00 = img + 4 * ((size_t)j0 * w + i0);
    const float *t10 = img + 4 * ((size_t)j0 * w + i1);
    const float *t01 = img + 4 * ((size_t)j1 * w + i0);
    const float *t11 = img + 4 * ((size_t)j1 * w + i1);
    for (int c = 0; c < 4; c++)
        out[c] = ((w00 * t00[c] + w10 * t10[c]) + w01 * t01[c]) + w11 * t11[c];
}

#endif /* FLX_MATH_H */
