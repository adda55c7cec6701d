// common.cuh — shared device helpers for libg2pc (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/g2pc.h"

// ---- error plumbing (no exceptions across the C ABI) ------------------------------------------------
void g2pc_set_error(const char* fmt, ...);

#define G2PC_CHECK_ARG(cond, msg)                                  \
    do {                                                           \
        if (!(cond)) {                                             \
            g2pc_set_error("%s: %s", __func__, msg);               \
            return G2PC_ERR_INVALID;                               \
        }                                                          \
    } while (0)

#define G2PC_CHECK_LAUNCH()                                                        \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            g2pc_set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e__)); \
            return G2PC_ERR_CUDA;                                                  \
        }                                                                          \
    } while (0)

#define G2PC_CUDA(call)                                                            \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            g2pc_set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e__)); \
            return G2PC_ERR_CUDA;                                                  \
        }                                                                          \
    } while (0)

// ---- Philox4x32-10 counter-based RNG -----------------------------------------------------------------
// Stream definition (sharding-invariant, regenerable in pass 2):
//   key = (seed_lo, seed_hi), counter = (global Gaussian id, sample index, attempt, call id)
// The same function is restated in oracle/philox.py.
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// u32 -> uniform in (0,1): ((x >> 9) + 0.5) * 2^-23.  x >> 9 < 2^23, so the +0.5 is exact in fp32 and the result lies
// in [2^-24, 1 - 2^-24]: never 0, never 1.
__device__ __forceinline__ float u32_to_unit(uint32_t x) {
    return ((float)(x >> 9) + 0.5f) * 1.1920928955078125e-07f;
}

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// three standard normals for (gid, sample, attempt, call): two Box-Muller pairs, the 4th value unused
__device__ __forceinline__ float3 draw_eps(uint32_t gid, uint32_t sample, uint32_t attempt, uint32_t call,
                                           uint32_t k0, uint32_t k1) {
    Philox4 r = philox4x32_10(gid, sample, attempt, call, k0, k1);
    const float TWO_PI = 6.283185307179586f;
    float u1 = u32_to_unit(r.x), u2 = u32_to_unit(r.y), u3 = u32_to_unit(r.z), u4 = u32_to_unit(r.w);
    // the approximate log of u just below 1 may come out slightly positive: clamp before the square root
    float ra = fast_sqrt(fmaxf(0.0f, -2.0f * __logf(u1)));
    float rb = fast_sqrt(fmaxf(0.0f, -2.0f * __logf(u3)));
    float sa, ca, cb;
    __sincosf(TWO_PI * u2, &sa, &ca);
    cb = __cosf(TWO_PI * u4);
    return make_float3(ra * ca, ra * sa, rb * cb);
}

// x = mu + L*eps with L lower-triangular (l00,l10,l11,l20,l21,l22); the SAME expression is used by the
// count pass (explicit Mahalanobis) and the emit pass, so both see bit-identical positions.
__device__ __forceinline__ float3 mvn_point(const float3 mu, float l00, float l10, float l11, float l20,
                                            float l21, float l22, const float3 e) {
    float dx = l00 * e.x;
    float dy = fmaf(l11, e.y, l10 * e.x);
    float dz = fmaf(l22, e.z, fmaf(l21, e.y, l20 * e.x));
    return make_float3(mu.x + dx, mu.y + dy, mu.z + dz);
}
