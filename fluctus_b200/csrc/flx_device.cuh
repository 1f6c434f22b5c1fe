// flx_device.cuh -- device-side vocabulary shared by all kernels of the wavefront path:
// 3-vectors with a pinned operation order, the path-state SoA accessors, the RNG and
// warp-aggregated queue pushes.
//
// Arithmetic contract (DESIGN.md "Numerics"): this translation unit is compiled with
// -fmad=false, IEEE division and square root, no flush-to-zero.  Every expression below is
// written in the association order the reference's OpenCL C source implies, so results are
// bit-identical to the oracle (reference kernels compiled for the host, oracle/_ref) which
// pins the same order in oracle/ref_shim/cl_shim.hpp.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "fluctus_b200.h"
#include "flx_math.h"

#define FLX_DEV __device__ __forceinline__

struct V3
{
    float x, y, z;
};

FLX_DEV V3 v3(float x, float y, float z) { return V3{x, y, z}; }
FLX_DEV V3 v3(float s) { return V3{s, s, s}; }
FLX_DEV V3 v3(const flx_float3 &f) { return V3{f.x, f.y, f.z}; }
FLX_DEV V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
FLX_DEV V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
FLX_DEV V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
FLX_DEV V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
FLX_DEV V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
FLX_DEV V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
FLX_DEV V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
FLX_DEV float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
FLX_DEV V3 cross3(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
FLX_DEV float len3(V3 a) { return sqrtf(dot3(a, a)); }
FLX_DEV V3 norm3(V3 a) // OpenCL normalize: normalize(0) = 0
{
    const float len = len3(a);
    if (len == 0.0f)
        return a;
    const float inv = 1.0f / len;
    return V3{a.x * inv, a.y * inv, a.z * inv};
}
FLX_DEV bool is_zero3(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
// (1-u-v)*a + u*b + v*c  (reference: src/utils.cl:27-30)
FLX_DEV V3 bary3(float u, float v, V3 a, V3 b, V3 c) { return ((1.0f - u - v) * a + u * b) + v * c; }

// ---- path-state SoA: slot s of path g at tasks[s * N + g]  (reference: src/geom.h:37-49)
struct Tasks
{
    uint32_t *base;
    uint32_t n;
    // Address of slot s of path g = (base + g) + s * n.  Written so that the per-path part (base + 4 g, two instructions, shared by
    // every access of a thread) is one common subexpression and the per-slot part is ONE IMAD.WIDE (n * (4 s) + that pointer): the
    // straightforward base[(size_t)s * n + g] costs three address instructions per access (IMAD.WIDE + LEA + LEA.HI.X), which made
    // path-state addressing a quarter of all instructions the fused logic kernel issues (profiles/r2_base_logic_lines.txt).
    FLX_DEV uint32_t *at(int slot, uint32_t g) const
    {
        char *p = reinterpret_cast<char *>(base) + (size_t)g * 4u;
        asm("" : "+l"(p)); // opaque to the optimiser: otherwise it re-associates to (n * 4s + 4g) + base = IMAD.WIDE + IADD3 + IADD3.X per access
        __builtin_assume(__isGlobal(p)); // (the empty asm hides where the pointer came from; without this the accesses become generic LD / ST)
        return reinterpret_cast<uint32_t *>(p + (size_t)n * (uint32_t)(slot * 4));
    }
    FLX_DEV uint32_t &u(int slot, uint32_t g) const { return *at(slot, g); }
    FLX_DEV float f(int slot, uint32_t g) const { return __uint_as_float(*at(slot, g)); }
    FLX_DEV void setf(int slot, uint32_t g, float v) const { *at(slot, g) = __float_as_uint(v); }
    FLX_DEV void setu(int slot, uint32_t g, uint32_t v) const { *at(slot, g) = v; }
    FLX_DEV V3 v(int slot, uint32_t g) const { return V3{f(slot, g), f(slot + 1, g), f(slot + 2, g)}; }
    // streaming flavours (ld/st.global.cs: evict-first): for state touched once by a kernel whose L1 is busy caching the BVH
    FLX_DEV float f_cs(int slot, uint32_t g) const { return __uint_as_float(__ldcs(at(slot, g))); }
    FLX_DEV uint32_t u_cs(int slot, uint32_t g) const { return __ldcs(at(slot, g)); }
    FLX_DEV V3 v_cs(int slot, uint32_t g) const { return V3{f_cs(slot, g), f_cs(slot + 1, g), f_cs(slot + 2, g)}; }
    FLX_DEV void setf_cs(int slot, uint32_t g, float v) const { __stcs(at(slot, g), __float_as_uint(v)); }
    FLX_DEV void setu_cs(int slot, uint32_t g, uint32_t v) const { __stcs(at(slot, g), v); }
    FLX_DEV void setv_cs(int slot, uint32_t g, V3 a) const
    {
        setf_cs(slot, g, a.x);
        setf_cs(slot + 1, g, a.y);
        setf_cs(slot + 2, g, a.z);
    }
    FLX_DEV void setv(int slot, uint32_t g, V3 a) const
    {
        setf(slot, g, a.x);
        setf(slot + 1, g, a.y);
        setf(slot + 2, g, a.z);
    }
};

// ---- RNG (reference: src/random.cl:7-22)
FLX_DEV uint32_t flx_hash(uint32_t s)
{
    s = (s ^ 61u) ^ (s >> 16);
    s *= 9u;
    s = s ^ (s >> 4);
    s *= 0x27d4eb2du;
    s = s ^ (s >> 15);
    return s;
}
FLX_DEV float flx_rand(uint32_t &seed)
{
    seed = flx_hash(seed);
    return (float)seed * (1.0f / 4294967296.0f);
}

// ---- warp-aggregated queue push: one atomic per warp per queue (reference intent: ptx_asm.cl:83-111,
// utils.cl:328-358).  Must be called by all lanes that are active at the call site with their own `pred`.
FLX_DEV uint32_t warp_push(uint32_t *counter, bool pred)
{
    const unsigned active = __activemask();
    const unsigned mask = __ballot_sync(active, pred);
    if (!pred)
        return 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader)
        base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}
