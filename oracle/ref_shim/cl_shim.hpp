// cl_shim.hpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// A host-side stand-in for the subset of OpenCL C 1.2 that the reference's wavefront kernels
// use, so that /root/reference/src/wf_*.cl (+ bvh.cl, intersect.cl, utils.cl, random.cl, the
// BSDF files and env_map.cl) can be compiled UNMODIFIED by g++ (after the single textual
// transform "(float3)(" -> "float3(" done by build_ref.py) and executed one work-item at a time.
// Nothing here restates the reference's algorithm: it only supplies the language run time.
//
// Pinned semantics (implementation-defined in OpenCL under -cl-fast-relaxed-math,
// reference src/clcontext.cpp:143-153):
//   * sin/cos/tan/atan2/acos/pow/native_sin/native_cos -> include/flx_math.h
//   * native_recip(x) = 1.0f/x (IEEE), sqrt = IEEE, dot = ((x*x')+(y*y'))+(z*z'),
//     length = sqrt(dot), normalize(v) = v * (1.0f/length(v)), normalize(0) = 0,
//   * fmin/fmax return the non-NaN operand (C99 / OpenCL 6.12.2),
//   * read_imagef(linear, normalized, clamp-to-edge) = flx_bilinear_rgba (OpenCL 1.2 sec. 8.2).
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

#include "flx_math.h"

typedef unsigned int uint;
typedef unsigned char uchar;

// ---------------------------------------------------------------- vector types
struct float2;
struct float3;
struct float4;

struct alignas(8) float2
{
    float x, y;
    float2() = default;
    float2(float s) : x(s), y(s) {}
    template <class A, class B> float2(A a, B b) : x((float)a), y((float)b) {}
};

struct xy_proxy
{
    float x, y;
    operator float2() const { return float2(x, y); }
};

struct xyz_proxy
{
    float x, y, z;
    inline operator float3() const;
    inline xyz_proxy &operator=(const float3 &v);
    xyz_proxy &operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct alignas(16) float3
{
    union
    {
        struct
        {
            float x, y, z, w;
        };
        xyz_proxy xyz;
        xy_proxy xy;
    };
    float3() = default;
    float3(float s) : x(s), y(s), z(s), w(0.0f) {}
    template <class A, class B, class C> float3(A a, B b, C c) : x((float)a), y((float)b), z((float)c), w(0.0f) {}
    inline float3(const float4 &v); // OpenCL allows float4 -> float3 only explicitly; the kernels do it via WriteFloat3 only
    float3 &operator+=(const float3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    float3 &operator-=(const float3 &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    float3 &operator*=(const float3 &o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
    float3 &operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
    float3 &operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};

struct alignas(16) float4
{
    union
    {
        struct
        {
            float x, y, z, w;
        };
        xyz_proxy xyz;
    };
    float4() = default;
    float4(float s) : x(s), y(s), z(s), w(s) {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float4(const float3 &v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
};

inline float3::float3(const float4 &v) : x(v.x), y(v.y), z(v.z), w(0.0f) {}
inline xyz_proxy::operator float3() const { return float3(x, y, z); }
inline xyz_proxy &xyz_proxy::operator=(const float3 &v) { x = v.x; y = v.y; z = v.z; return *this; }

struct alignas(8) int2
{
    int x, y;
    int2() = default;
    int2(int s) : x(s), y(s) {}
    template <class A, class B> int2(A a, B b) : x((int)a), y((int)b) {}
};

static_assert(sizeof(float3) == 16 && sizeof(float4) == 16 && sizeof(float2) == 8, "vector sizes");

inline float3 operator+(const float3 &a, const float3 &b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(const float3 &a, const float3 &b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator*(const float3 &a, const float3 &b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator/(const float3 &a, const float3 &b) { return float3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline float3 operator*(const float3 &a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float s, const float3 &a) { return float3(s * a.x, s * a.y, s * a.z); }
inline float3 operator/(const float3 &a, float s) { return float3(a.x / s, a.y / s, a.z / s); }
inline float3 operator/(float s, const float3 &a) { return float3(s / a.x, s / a.y, s / a.z); }
inline float3 operator+(const float3 &a, float s) { return float3(a.x + s, a.y + s, a.z + s); }
inline float3 operator+(float s, const float3 &a) { return float3(s + a.x, s + a.y, s + a.z); }
inline float3 operator-(const float3 &a, float s) { return float3(a.x - s, a.y - s, a.z - s); }
inline float3 operator-(float s, const float3 &a) { return float3(s - a.x, s - a.y, s - a.z); }
inline float3 operator-(const float3 &a) { return float3(-a.x, -a.y, -a.z); }
inline float4 operator+(const float4 &a, const float4 &b) { return float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline float4 &operator+=(float4 &a, const float4 &b) { a = a + b; return a; } // mk_splat.cl:22
inline float4 operator*(const float4 &a, float s) { return float4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline float4 operator/(const float4 &a, float s) { return float4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline float2 operator*(const float2 &a, float s) { return float2(a.x * s, a.y * s); }
inline float2 operator*(float s, const float2 &a) { return float2(s * a.x, s * a.y); }
inline float2 operator+(const float2 &a, const float2 &b) { return float2(a.x + b.x, a.y + b.y); }

// ---------------------------------------------------------------- built-ins
inline float cl_dot(const float3 &a, const float3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float cl_dot(const float4 &a, const float4 &b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline float3 cl_cross(const float3 &a, const float3 &b)
{
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float cl_length(const float3 &a) { return sqrtf(cl_dot(a, a)); }
inline float3 cl_normalize(const float3 &a)
{
    const float len = cl_length(a);
    if (len == 0.0f)
        return a;
    const float inv = 1.0f / len;
    return float3(a.x * inv, a.y * inv, a.z * inv);
}
inline float cl_fmin(float a, float b) { return fminf(a, b); }
inline float cl_fmax(float a, float b) { return fmaxf(a, b); }
inline float3 cl_fmin(const float3 &a, const float3 &b) { return float3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
inline float3 cl_fmax(const float3 &a, const float3 &b) { return float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
inline float cl_sqrt(float a) { return sqrtf(a); }
inline float3 cl_sqrt(const float3 &a) { return float3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }
inline float cl_fabs(float a) { return fabsf(a); }
inline float cl_floor(float a) { return floorf(a); }
#ifdef SHIM_LIBM
// The "lm_" build (oracle/build_ref.py): the C library's single-precision functions instead of the shared include/flx_math.h.
// OpenCL leaves these built-ins implementation-defined (under -cl-fast-relaxed-math even more so), so glibc is as legitimate a
// stand-in as flx_math.h -- and an INDEPENDENT one: GPU, C restatement and the base oracle all share flx_math.h, so a bias in
// that header is invisible to every bit-parity test.  A converged image of this build against the GPU's is what can see it.
#define SHIM_SIN(a) sinf(a)
#define SHIM_COS(a) cosf(a)
#define SHIM_TAN(a) tanf(a)
#define SHIM_ACOS(a) acosf(a)
#define SHIM_ATAN2(y, x) atan2f(y, x)
#define SHIM_POW(a, b) powf(a, b)
#else
#define SHIM_SIN(a) flx_sinf(a)
#define SHIM_COS(a) flx_cosf(a)
#define SHIM_TAN(a) flx_tanf(a)
#define SHIM_ACOS(a) flx_acosf(a)
#define SHIM_ATAN2(y, x) flx_atan2f(y, x)
#define SHIM_POW(a, b) flx_powf(a, b)
#endif
inline float cl_sin(float a) { return SHIM_SIN(a); }
inline float cl_cos(float a) { return SHIM_COS(a); }
inline float cl_tan(float a) { return SHIM_TAN(a); }
inline float cl_acos(float a) { return SHIM_ACOS(a); }
inline float cl_atan2(float y, float x) { return SHIM_ATAN2(y, x); }
inline float cl_pow(float a, float b) { return SHIM_POW(a, b); }
inline float3 cl_pow(const float3 &a, float b) { return float3(SHIM_POW(a.x, b), SHIM_POW(a.y, b), SHIM_POW(a.z, b)); }
inline float native_recip(float a) { return 1.0f / a; }
inline float3 native_recip(const float3 &a) { return float3(1.0f / a.x, 1.0f / a.y, 1.0f / a.z); }
inline float native_sin(float a) { return SHIM_SIN(a); }
inline float native_cos(float a) { return SHIM_COS(a); }
inline float native_powr(float a, float b) { return SHIM_POW(a, b); }

template <class T> inline T cl_max(T a, T b) { return a < b ? b : a; }
template <class T> inline T cl_min(T a, T b) { return b < a ? b : a; }
inline float cl_max(float a, float b) { return fmaxf(a, b); }
inline float cl_min(float a, float b) { return fminf(a, b); }
inline uint cl_min(uint a, uint b) { return b < a ? b : a; }
inline uint cl_max(uint a, uint b) { return a < b ? b : a; }
inline int cl_min(int a, int b) { return b < a ? b : a; }
inline int cl_max(int a, int b) { return a < b ? b : a; }
inline float cl_clamp(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
inline int2 cl_clamp(const int2 &v, const int2 &lo, const int2 &hi)
{
    return int2(cl_min(cl_max(v.x, lo.x), hi.x), cl_min(cl_max(v.y, lo.y), hi.y));
}

inline float4 vload4(size_t off, const float *p) { return float4(p[4 * off + 0], p[4 * off + 1], p[4 * off + 2], p[4 * off + 3]); }
inline void vstore4(const float4 &v, size_t off, float *p)
{
    p[4 * off + 0] = v.x; p[4 * off + 1] = v.y; p[4 * off + 2] = v.z; p[4 * off + 3] = v.w;
}

// ---------------------------------------------------------------- work-item ids, atomics
static thread_local size_t g_shim_gid = 0;
static thread_local size_t g_shim_gid1 = 0; // second NDRange dimension: only the microkernel reset/splat kernels are launched 2-D (clcontext.cpp:712,742,748)
inline size_t get_global_id(int dim) { return dim == 1 ? g_shim_gid1 : g_shim_gid; }
inline size_t get_local_id(int) { return g_shim_gid % 32; }

#ifdef SHIM_PARALLEL
inline uint atomic_inc(volatile uint *p) { return __atomic_fetch_add((uint *)p, 1u, __ATOMIC_RELAXED); }
inline uint atomic_add(volatile uint *p, uint v) { return __atomic_fetch_add((uint *)p, v, __ATOMIC_RELAXED); }
inline float atomic_xchg(volatile float *p, float v)
{
    uint in, out;
    memcpy(&in, &v, 4);
    out = __atomic_exchange_n((uint *)p, in, __ATOMIC_RELAXED);
    float r;
    memcpy(&r, &out, 4);
    return r;
}
inline uint atomic_cmpxchg(volatile uint *p, uint cmp, uint val)
{
    __atomic_compare_exchange_n((uint *)p, &cmp, val, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return cmp;
}
#else
inline uint atomic_inc(volatile uint *p) { uint o = *p; *p = o + 1; return o; }
inline uint atomic_add(volatile uint *p, uint v) { uint o = *p; *p = o + v; return o; }
inline float atomic_xchg(volatile float *p, float v) { float o = *p; *p = v; return o; }
inline uint atomic_cmpxchg(volatile uint *p, uint cmp, uint val) { uint o = *p; if (o == cmp) *p = val; return o; }
#endif

#define CLK_LOCAL_MEM_FENCE 1
inline void barrier(int) {}

// ---------------------------------------------------------------- images
struct shim_image
{
    int width, height;
    const float *rgba;
};
typedef const shim_image *image2d_t;
typedef uint sampler_t;
#define CLK_NORMALIZED_COORDS_FALSE 0u
#define CLK_NORMALIZED_COORDS_TRUE 1u
#define CLK_ADDRESS_CLAMP_TO_EDGE 2u
#define CLK_FILTER_NEAREST 0u
#define CLK_FILTER_LINEAR 4u

inline int2 get_image_dim(image2d_t img) { return int2(img->width, img->height); }
inline float4 read_imagef(image2d_t img, sampler_t, const float2 &uv)
{
    float o[4];
    flx_bilinear_rgba(img->rgba, img->width, img->height, uv.x, uv.y, o);
    return float4(o[0], o[1], o[2], o[3]);
}
inline float4 read_imagef(image2d_t img, sampler_t, const int2 &c)
{
    const int x = cl_min(cl_max(c.x, 0), img->width - 1), y = cl_min(cl_max(c.y, 0), img->height - 1);
    const float *p = img->rgba + 4 * ((size_t)y * img->width + x);
    return float4(p[0], p[1], p[2], p[3]);
}

// ---------------------------------------------------------------- keywords and names (must come last)
#define M_PI_F FLX_PI_F
#define global
#define __global
#define kernel
#define constant const
#define read_only
#define dot cl_dot
#define cross cl_cross
#define length cl_length
#define normalize cl_normalize
#define fmin cl_fmin
#define fmax cl_fmax
#define sqrt cl_sqrt
#define fabs cl_fabs
#define floor cl_floor
#define sin cl_sin
#define cos cl_cos
#define tan cl_tan
#define acos cl_acos
#define atan2 cl_atan2
#define pow cl_pow
#define max cl_max
#define min cl_min
#define clamp cl_clamp
