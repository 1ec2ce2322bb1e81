// Shared device/host helpers for the plnlp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <initializer_list>

#include "../../include/plnlp_b200.h"

namespace plnlp {

extern long long g_launch_count;  // defined in abi.cu

#define PLNLP_REQUIRE(cond, code) \
    do {                          \
        if (!(cond)) return (code); \
    } while (0)

// Call after every launch: counts it and converts a launch failure into the ABI's
// positive-cudaError return.
#define PLNLP_LAUNCH_CHECK()                                    \
    do {                                                        \
        ++::plnlp::g_launch_count;                              \
        cudaError_t _e = cudaGetLastError();                    \
        if (_e != cudaSuccess) return static_cast<int>(_e);     \
    } while (0)

constexpr int kNumSM = 148;

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }


// ---------------------------------------------------------------------------------------
// VEC-wide (1/2/4 floats; 8 = two 16-byte accesses) global loads and stores
// ---------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void load_vec(float (&d)[VEC], const float* p) {
    static_assert(VEC == 1 || VEC == 2 || VEC == 4 || VEC == 8, "unsupported vector width");
    if constexpr (VEC == 8) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
    } else if constexpr (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
    } else if constexpr (VEC == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        d[0] = t.x; d[1] = t.y;
    } else {
        d[0] = __ldg(p);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&d)[VEC]) {
    static_assert(VEC == 1 || VEC == 2 || VEC == 4 || VEC == 8, "unsupported vector width");
    if constexpr (VEC == 8) {
        reinterpret_cast<float4*>(p)[0] = make_float4(d[0], d[1], d[2], d[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(d[4], d[5], d[6], d[7]);
    } else if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(d[0], d[1]);
    } else {
        p[0] = d[0];
    }
}

// widest legal vector width for rows of `width` floats with the given leading dims / bases
inline int pick_vec(int64_t width, std::initializer_list<int64_t> lds, std::initializer_list<const void*> ptrs) {
    int v = 4;
    if (width % 4) v = (width % 2) ? 1 : 2;
    for (int64_t ld : lds) { if (ld % 4 && v == 4) v = 2; if (ld % 2 && v == 2) v = 1; }
    for (const void* q : ptrs) {
        if (!q) continue;
        if (!aligned(q, 16) && v == 4) v = 2;
        if (!aligned(q, 8) && v == 2) v = 1;
    }
    return v;
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), written out here; counter = (idx_lo, idx_hi, stream, 0),
// key = (seed_lo, seed_hi).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint64_t seed, uint64_t idx, uint32_t stream_id) {
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
    uint32_t c0 = static_cast<uint32_t>(idx), c1 = static_cast<uint32_t>(idx >> 32), c2 = stream_id, c3 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float u32_to_unit(uint32_t r) {  // [0, 1)
    return static_cast<float>(r >> 8) * (1.0f / 16777216.0f);
}

// Inverted-dropout keep decision for logical element `e` of a tensor: element e uses word
// (e & 3) of philox(seed, e >> 2).  One definition shared by every epilogue.
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t e, float p) {
    const uint4 r = philox4x32_10(seed, e >> 2, 0x0d0u);
    const uint32_t w = (e & 3) == 0 ? r.x : (e & 3) == 1 ? r.y : (e & 3) == 2 ? r.z : r.w;
    return u32_to_unit(w) >= p;
}

// keep decisions for 4 consecutive elements starting at e (e % 4 == 0)
__device__ __forceinline__ void dropout_keep4(uint64_t seed, uint64_t e, float p, bool k[4]) {
    const uint4 r = philox4x32_10(seed, e >> 2, 0x0d0u);
    k[0] = u32_to_unit(r.x) >= p; k[1] = u32_to_unit(r.y) >= p;
    k[2] = u32_to_unit(r.z) >= p; k[3] = u32_to_unit(r.w) >= p;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// order-preserving map float -> uint32 (larger float <=> larger key); -0 < +0, NaNs sort high.
__device__ __forceinline__ uint32_t float_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

}  // namespace plnlp
