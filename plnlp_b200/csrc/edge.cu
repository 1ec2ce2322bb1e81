// Edge scoring building blocks: endpoint gather (+Hadamard / +dot), the out_channels=1 MLP layer,
// and the backward of the endpoint gather (scatter-add into grad_h).
//
// Replaces h[edge[0]] / h[edge[1]] advanced indexing + MLPPredictor / DotPredictor
// (/root/reference/plnlp/model.py:152-156,180; plnlp/layer.py:80-87,174-176) and the
// index_put_(accumulate=True) backward autograd derives (model.py:161).
// All kernels are gather/stream bound: one warp per pair (or per node segment), lanes stride over
// the feature row with the widest legal vector width.
#include "common.cuh"

namespace plnlp {

__device__ __forceinline__ int64_t wrap_index(int64_t i, int64_t n) { return i < 0 ? i + n : i; }

// ---------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) gather_hadamard_kernel(const float* __restrict__ h, int64_t ldh,
                                                              int64_t n_rows, const int64_t* __restrict__ edges,
                                                              int64_t P, int H, float* __restrict__ out,
                                                              int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t p = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const int64_t s = wrap_index(__ldg(edges + 2 * p), n_rows), d = wrap_index(__ldg(edges + 2 * p + 1), n_rows);
    const float* hs = h + s * ldh;
    const float* hd = h + d * ldh;
    float* o = out + p * ldo;
    for (int f = lane * VEC; f < H; f += 32 * VEC) {
        float a[VEC], b[VEC];
        load_vec<VEC>(a, hs + f);
        load_vec<VEC>(b, hd + f);
#pragma unroll
        for (int e = 0; e < VEC; ++e) a[e] *= b[e];
        store_vec<VEC>(o + f, a);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) edge_dot_kernel(const float* __restrict__ h, int64_t ldh, int64_t n_rows,
                                                       const int64_t* __restrict__ edges, int64_t P, int H,
                                                       float* __restrict__ score) {
    const int lane = threadIdx.x & 31;
    const int64_t p = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const int64_t s = wrap_index(__ldg(edges + 2 * p), n_rows), d = wrap_index(__ldg(edges + 2 * p + 1), n_rows);
    const float* hs = h + s * ldh;
    const float* hd = h + d * ldh;
    float acc = 0.0f;
    for (int f = lane * VEC; f < H; f += 32 * VEC) {
        float a[VEC], b[VEC];
        load_vec<VEC>(a, hs + f);
        load_vec<VEC>(b, hd + f);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc = fmaf(a[e], b[e], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) score[p] = acc;
}

template <int VEC>
__global__ void __launch_bounds__(256) mlp_out_fwd_kernel(const float* __restrict__ a, int64_t lda,
                                                          const float* __restrict__ w, const float* __restrict__ b,
                                                          int64_t P, int H, float* __restrict__ score) {
    const int lane = threadIdx.x & 31;
    const int64_t p = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const float* ar = a + p * lda;
    float acc = 0.0f;
    for (int f = lane * VEC; f < H; f += 32 * VEC) {
        float x[VEC], y[VEC];
        load_vec<VEC>(x, ar + f);
        load_vec<VEC>(y, w + f);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc = fmaf(x[e], y[e], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) score[p] = acc + (b ? __ldg(b) : 0.0f);
}

// dz = dscore * w * mask(a);  per-block partial dw over MLP_RB consecutive rows (fixed order)
constexpr int MLP_RB = 128;

template <int VEC>
__global__ void __launch_bounds__(128) mlp_out_bwd_kernel(const float* __restrict__ a, int64_t lda,
                                                          const float* __restrict__ w,
                                                          const float* __restrict__ dscore, int64_t P, int H,
                                                          int mask_a, float drop_scale, float* __restrict__ dz,
                                                          int64_t lddz, float* __restrict__ ws_dw,
                                                          float* __restrict__ ws_dzsum) {
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * MLP_RB;
    const int64_t r1 = min(P, r0 + MLP_RB);
    __shared__ float ds[MLP_RB];
    for (int i = threadIdx.x; i < MLP_RB; i += blockDim.x) ds[i] = (r0 + i < P) ? __ldg(dscore + r0 + i) : 0.0f;
    __syncthreads();
    for (int f = threadIdx.x * VEC; f < H; f += blockDim.x * VEC) {
        float wv[VEC], acc[VEC], gsum[VEC];
        load_vec<VEC>(wv, w + f);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = gsum[e] = 0.0f;
        // rows are fetched RU at a time (independent loads in flight) and consumed in row order, so the
        // per-column sums keep their order
        constexpr int RU = 4;
        int64_t r = r0;
        for (; r + RU - 1 < r1; r += RU) {
            float av[RU][VEC];
#pragma unroll
            for (int i = 0; i < RU; ++i) load_vec<VEC>(av[i], a + (r + i) * lda + f);
#pragma unroll
            for (int i = 0; i < RU; ++i) {
                float g[VEC];
                const float d = ds[r + i - r0];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    acc[e] = fmaf(d, av[i][e], acc[e]);
                    g[e] = d * wv[e];
                    if (mask_a) g[e] = av[i][e] > 0.0f ? g[e] * drop_scale : 0.0f;
                    gsum[e] += g[e];
                }
                if (dz) store_vec<VEC>(dz + (r + i) * lddz + f, g);
            }
        }
        for (; r < r1; ++r) {
            float av[VEC], g[VEC];
            load_vec<VEC>(av, a + r * lda + f);
            const float d = ds[r - r0];
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                acc[e] = fmaf(d, av[e], acc[e]);
                g[e] = d * wv[e];
                if (mask_a) g[e] = av[e] > 0.0f ? g[e] * drop_scale : 0.0f;
                gsum[e] += g[e];
            }
            if (dz) store_vec<VEC>(dz + r * lddz + f, g);
        }
        store_vec<VEC>(ws_dw + static_cast<int64_t>(blockIdx.x) * H + f, acc);
        if (ws_dzsum) store_vec<VEC>(ws_dzsum + static_cast<int64_t>(blockIdx.x) * H + f, gsum);
    }
}

// dw[j] = sum over row blocks (in block order); the extra last block reduces db = sum dscore.
__global__ void __launch_bounds__(256) mlp_out_bwd_reduce_kernel(const float* __restrict__ ws_dw, int64_t nblk,
                                                                 int H, const float* __restrict__ dscore, int64_t P,
                                                                 float* __restrict__ dw, float* __restrict__ db,
                                                                 const float* __restrict__ ws_dzsum,
                                                                 float* __restrict__ dzsum) {
    const int ncolblk = (H + 255) / 256;
    if (static_cast<int>(blockIdx.x) < ncolblk) {
        const int j = blockIdx.x * 256 + threadIdx.x;
        if (j >= H) return;
        float acc = 0.0f;
        for (int64_t b = 0; b < nblk; ++b) acc += ws_dw[b * H + j];
        dw[j] = acc;
        if (dzsum) {                 // column sums of dz (the previous layer's bias gradient), fp64 over the blocks
            double s = 0.0;
            for (int64_t b = 0; b < nblk; ++b) s += static_cast<double>(ws_dzsum[b * H + j]);
            dzsum[j] = static_cast<float>(s);
        }
    } else {
        __shared__ double sm[256];
        double acc = 0.0;
        for (int64_t i = threadIdx.x; i < P; i += 256) acc += static_cast<double>(dscore[i]);
        sm[threadIdx.x] = acc;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (static_cast<int>(threadIdx.x) < o) sm[threadIdx.x] += sm[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0 && db) db[0] = static_cast<float>(sm[0]);
    }
}

// ---------------------------------------------------------------------------------------
// backward of the endpoint gather
// ---------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void red_add_vec(float* p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                     "f"(v[3])
                     : "memory");
    } else if constexpr (VEC == 2) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
        atomicAdd(p, v[0]);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) edge_scatter_atomic_kernel(const float* __restrict__ h, int64_t ldh,
                                                                  int64_t n_rows, const int64_t* __restrict__ edges,
                                                                  int64_t P, int H, const float* __restrict__ da,
                                                                  int64_t ldda, const float* __restrict__ dscore,
                                                                  float* __restrict__ grad_h, int64_t ldg) {
    const int lane = threadIdx.x & 31;
    const int64_t p = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const int64_t s = wrap_index(__ldg(edges + 2 * p), n_rows), d = wrap_index(__ldg(edges + 2 * p + 1), n_rows);
    const float gs = da ? 1.0f : __ldg(dscore + p);
    for (int f = lane * VEC; f < H; f += 32 * VEC) {
        float g[VEC], a[VEC], b[VEC];
        if (da) {
            load_vec<VEC>(g, da + p * ldda + f);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) g[e] = gs;
        }
        load_vec<VEC>(a, h + s * ldh + f);
        load_vec<VEC>(b, h + d * ldh + f);
        float us[VEC], ud[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            us[e] = g[e] * b[e];
            ud[e] = g[e] * a[e];
        }
        red_add_vec<VEC>(grad_h + s * ldg + f, us);
        red_add_vec<VEC>(grad_h + d * ldg + f, ud);
    }
}

// one warp per node segment; entries (2*p + side) of a segment are visited in list order
template <int VEC>
__global__ void __launch_bounds__(256) edge_scatter_sorted_kernel(const float* __restrict__ h, int64_t ldh,
                                                                  int64_t n_rows, const int64_t* __restrict__ edges,
                                                                  int H, const float* __restrict__ da, int64_t ldda,
                                                                  const float* __restrict__ dscore,
                                                                  const int64_t* __restrict__ seg_ptr,
                                                                  const int64_t* __restrict__ seg_node, int64_t n_seg,
                                                                  const int64_t* __restrict__ entry,
                                                                  float* __restrict__ grad_h, int64_t ldg) {
    // seg_node == NULL: segment sg IS node sg (one segment per node, empty ones write a zero row).
    // The (entry, pair, partner) triples are fetched 32 at a time, one per lane, and broadcast by
    // shuffle; each lane owns U feature vectors so every gathered row is read exactly once.
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int64_t sg = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (sg >= n_seg) return;
    const int64_t t0 = __ldg(seg_ptr + sg), t1 = __ldg(seg_ptr + sg + 1);
    const int64_t node = seg_node ? wrap_index(__ldg(seg_node + sg), n_rows) : sg;
    for (int f0 = 0; f0 < H; f0 += 32 * VEC * U) {
        float acc[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[u][e] = 0.0f;
        for (int64_t base = t0; base < t1; base += 32) {
            const int n = (t1 - base) < 32 ? static_cast<int>(t1 - base) : 32;
            int64_t p = 0, partner = 0;
            float gs = 0.0f;
            if (lane < n) {
                const int64_t ent = __ldg(entry + base + lane);
                p = ent >> 1;
                partner = wrap_index(__ldg(edges + 2 * p + ((ent & 1) ^ 1)), n_rows);
                if (!da) gs = __ldg(dscore + p);
            }
            for (int j = 0; j < n; ++j) {
                const int64_t pj = __shfl_sync(0xffffffffu, p, j);
                const int64_t qj = __shfl_sync(0xffffffffu, partner, j);
                const float gj = __shfl_sync(0xffffffffu, gs, j);
                float g[U][VEC], b[U][VEC];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int f = f0 + (u * 32 + lane) * VEC;
                    if (f < H) {
                        if (da) {
                            load_vec<VEC>(g[u], da + pj * ldda + f);
                        } else {
#pragma unroll
                            for (int e = 0; e < VEC; ++e) g[u][e] = gj;
                        }
                        load_vec<VEC>(b[u], h + qj * ldh + f);
                    } else {
#pragma unroll
                        for (int e = 0; e < VEC; ++e) g[u][e] = b[u][e] = 0.0f;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[u][e] = fmaf(g[u][e], b[u][e], acc[u][e]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int f = f0 + (u * 32 + lane) * VEC;
            if (f < H) store_vec<VEC>(grad_h + node * ldg + f, acc[u]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// plain row gather out[p,:] = h[idx[p*idx_stride],:] and its backward (sorted segment sum): the pair-level
// predictors of layer.py:90-189 (MLPCAT, and MLPDOT / MLPBIL while dropout is active) need the two endpoint
// rows separately rather than their product.
// ---------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ h, int64_t ldh, int64_t n_rows,
                                                          const int64_t* __restrict__ idx, int64_t idx_stride,
                                                          int64_t P, int H, float* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t p = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const float* src = h + wrap_index(__ldg(idx + p * idx_stride), n_rows) * ldh;
    float* o = out + p * ldo;
    for (int f = lane * VEC; f < H; f += 32 * VEC) {
        float a[VEC];
        load_vec<VEC>(a, src + f);
        store_vec<VEC>(o + f, a);
    }
}

// one warp per node: grad_h[node,:] = sum of g[entry[t],:] over the node's segment, in list order
template <int VEC>
__global__ void __launch_bounds__(256) row_scatter_sorted_kernel(const float* __restrict__ g, int64_t ldg_in, int H,
                                                                 const int64_t* __restrict__ seg_ptr, int64_t n_seg,
                                                                 const int64_t* __restrict__ entry,
                                                                 float* __restrict__ grad_h, int64_t ldg) {
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int64_t sg = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (sg >= n_seg) return;
    const int64_t t0 = __ldg(seg_ptr + sg), t1 = __ldg(seg_ptr + sg + 1);
    for (int f0 = 0; f0 < H; f0 += 32 * VEC * U) {
        float acc[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[u][e] = 0.0f;
        for (int64_t base = t0; base < t1; base += 32) {
            const int n = (t1 - base) < 32 ? static_cast<int>(t1 - base) : 32;
            int64_t p = 0;
            if (lane < n) p = __ldg(entry + base + lane);
            for (int j = 0; j < n; ++j) {
                const int64_t pj = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int f = f0 + (u * 32 + lane) * VEC;
                    if (f < H) {
                        float v[VEC];
                        load_vec<VEC>(v, g + pj * ldg_in + f);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[u][e] += v[e];
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int f = f0 + (u * 32 + lane) * VEC;
            if (f < H) store_vec<VEC>(grad_h + sg * ldg + f, acc[u]);
        }
    }
}

}  // namespace plnlp

using namespace plnlp;

#define DISPATCH_VEC(vec, CALL)      \
    do {                             \
        if ((vec) == 4) { CALL(4); } \
        else if ((vec) == 2) { CALL(2); } \
        else { CALL(1); }            \
    } while (0)

extern "C" int plnlp_gather_hadamard_f32(const float* h, int64_t ldh, int64_t n_rows, const int64_t* edges,
                                         int64_t P, int64_t H, float* out, int64_t ldo, void* stream) {
    PLNLP_REQUIRE(P >= 0 && H > 0 && n_rows > 0, PLNLP_E_SIZE);
    if (P == 0) return 0;
    PLNLP_REQUIRE(h && edges && out, PLNLP_E_NULL);
    PLNLP_REQUIRE(ldh >= H && ldo >= H, PLNLP_E_SIZE);
    const int vec = pick_vec(H, {ldh, ldo}, {h, out});
    const unsigned grid = static_cast<unsigned>(ceil_div(P, 8));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(V) gather_hadamard_kernel<V><<<grid, 256, 0, st>>>(h, ldh, n_rows, edges, P, static_cast<int>(H), out, ldo)
    DISPATCH_VEC(vec, CALL);
#undef CALL
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int plnlp_edge_dot_fwd_f32(const float* h, int64_t ldh, int64_t n_rows, const int64_t* edges, int64_t P,
                                      int64_t H, float* score, void* stream) {
    PLNLP_REQUIRE(P >= 0 && H > 0 && n_rows > 0, PLNLP_E_SIZE);
    if (P == 0) return 0;
    PLNLP_REQUIRE(h && edges && score, PLNLP_E_NULL);
    PLNLP_REQUIRE(ldh >= H, PLNLP_E_SIZE);
    const int vec = pick_vec(H, {ldh}, {h});
    const unsigned grid = static_cast<unsigned>(ceil_div(P, 8));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(V) edge_dot_kernel<V><<<grid, 256, 0, st>>>(h, ldh, n_rows, edges, P, static_cast<int>(H), score)
    DISPATCH_VEC(vec, CALL);
#undef CALL
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int plnlp_mlp_out_fwd_f32(const float* a, int64_t lda, const float* w, const float* b, int64_t P,
                                     int64_t H, float* score, void* stream) {
    PLNLP_REQUIRE(P >= 0 && H > 0, PLNLP_E_SIZE);
    if (P == 0) return 0;
    PLNLP_REQUIRE(a && w && score, PLNLP_E_NULL);
    PLNLP_REQUIRE(lda >= H, PLNLP_E_SIZE);
    const int vec = pick_vec(H, {lda}, {a, w});
    const unsigned grid = static_cast<unsigned>(ceil_div(P, 8));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(V) mlp_out_fwd_kernel<V><<<grid, 256, 0, st>>>(a, lda, w, b, P, static_cast<int>(H), score)
    DISPATCH_VEC(vec, CALL);
#undef CALL
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t plnlp_mlp_out_bwd_workspace_bytes(int64_t P, int64_t H) {
    if (P < 0 || H < 0) return 0;
    return 2 * ceil_div(P, MLP_RB) * H * 4 + 16;
}

extern "C" int plnlp_mlp_out_bwd_f32(const float* a, int64_t lda, const float* w, const float* dscore, int64_t P,
                                     int64_t H, int mask_a, float drop_scale, float* dz, int64_t lddz, float* dw,
                                     float* db, float* dzsum, void* workspace, int64_t workspace_bytes,
                                     void* stream) {
    PLNLP_REQUIRE(P >= 0 && H > 0, PLNLP_E_SIZE);
    PLNLP_REQUIRE(a && w && dscore && dw && workspace, PLNLP_E_NULL);
    PLNLP_REQUIRE(lda >= H && (!dz || lddz >= H), PLNLP_E_SIZE);
    PLNLP_REQUIRE(workspace_bytes >= plnlp_mlp_out_bwd_workspace_bytes(P, H), PLNLP_E_WORKSPACE);
    PLNLP_REQUIRE(aligned(workspace, 16), PLNLP_E_ALIGN);
    const int64_t nblk = ceil_div(P, MLP_RB);
    float* ws = static_cast<float*>(workspace);
    float* ws2 = dzsum ? ws + nblk * H : nullptr;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (nblk > 0) {
        const int vec = pick_vec(H, {lda, dz ? lddz : 4}, {a, w, dz, ws2});
        const unsigned grid = static_cast<unsigned>(nblk);
#define CALL(V) \
    mlp_out_bwd_kernel<V><<<grid, 128, 0, st>>>(a, lda, w, dscore, P, static_cast<int>(H), mask_a, drop_scale, dz, lddz, ws, ws2)
        DISPATCH_VEC(vec, CALL);
#undef CALL
        PLNLP_LAUNCH_CHECK();
    }
    const unsigned rgrid = static_cast<unsigned>((H + 255) / 256 + 1);
    mlp_out_bwd_reduce_kernel<<<rgrid, 256, 0, st>>>(ws, nblk, static_cast<int>(H), dscore, P, dw, db, ws2, dzsum);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int plnlp_edge_scatter_atomic_f32(const float* h, int64_t ldh, int64_t n_rows, const int64_t* edges,
                                             int64_t P, int64_t H, const float* da, int64_t ldda,
                                             const float* dscore, float* grad_h, int64_t ldg, void* stream) {
    PLNLP_REQUIRE(P >= 0 && H > 0 && n_rows > 0, PLNLP_E_SIZE);
    if (P == 0) return 0;
    PLNLP_REQUIRE(h && edges && grad_h && (da || dscore), PLNLP_E_NULL);
    PLNLP_REQUIRE(ldh >= H && ldg >= H && (!da || ldda >= H), PLNLP_E_SIZE);
    const int vec = pick_vec(H, {ldh, ldg, da ? ldda : 4}, {h, grad_h, da});
    const unsigned grid = static_cast<unsigned>(ceil_div(P, 8));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(V) \
    edge_scatter_atomic_kernel<V><<<grid, 256, 0, st>>>(h, ldh, n_rows, edges, P, static_cast<int>(H), da, ldda, dscore, grad_h, ldg)
    DISPATCH_VEC(vec, CALL);
#undef CALL
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int plnlp_edge_scatter_sorted_f32(const float* h, int64_t ldh, int64_t n_rows, const int64_t* edges,
                                             int64_t P, int64_t H, const float* da, int64_t ldda,
                                             const float* dscore, const int64_t* seg_ptr, const int64_t* seg_node,
                                             int64_t n_seg, const int64_t* entry, float* grad_h, int64_t ldg,
                                             void* stream) {
    PLNLP_REQUIRE(P >= 0 && H > 0 && n_rows > 0 && n_seg >= 0, PLNLP_E_SIZE);
    if (n_seg == 0) return 0;
    PLNLP_REQUIRE(h && edges && grad_h && (da || dscore) && seg_ptr && entry, PLNLP_E_NULL);
    PLNLP_REQUIRE(seg_node || n_seg <= n_rows, PLNLP_E_SIZE);
    PLNLP_REQUIRE(ldh >= H && ldg >= H && (!da || ldda >= H), PLNLP_E_SIZE);
    const int vec = pick_vec(H, {ldh, ldg, da ? ldda : 4}, {h, grad_h, da});
    const unsigned grid = static_cast<unsigned>(ceil_div(n_seg, 8));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(V) \
    edge_scatter_sorted_kernel<V><<<grid, 256, 0, st>>>(h, ldh, n_rows, edges, static_cast<int>(H), da, ldda, dscore, seg_ptr, seg_node, n_seg, entry, grad_h, ldg)
    DISPATCH_VEC(vec, CALL);
#undef CALL
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int plnlp_gather_rows_f32(const float* h, int64_t ldh, int64_t n_rows, const int64_t* idx,
                                     int64_t idx_stride, int64_t P, int64_t H, float* out, int64_t ldo,
                                     void* stream) {
    PLNLP_REQUIRE(P >= 0 && H > 0 && n_rows > 0 && idx_stride >= 1, PLNLP_E_SIZE);
    if (P == 0) return 0;
    PLNLP_REQUIRE(h && idx && out, PLNLP_E_NULL);
    PLNLP_REQUIRE(ldh >= H && ldo >= H, PLNLP_E_SIZE);
    const int vec = pick_vec(H, {ldh, ldo}, {h, out});
    const unsigned grid = static_cast<unsigned>(ceil_div(P, 8));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(V) \
    gather_rows_kernel<V><<<grid, 256, 0, st>>>(h, ldh, n_rows, idx, idx_stride, P, static_cast<int>(H), out, ldo)
    DISPATCH_VEC(vec, CALL);
#undef CALL
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int plnlp_row_scatter_sorted_f32(const float* g, int64_t ldg_in, int64_t H, const int64_t* seg_ptr,
                                            int64_t n_seg, const int64_t* entry, float* grad_h, int64_t ldg,
                                            void* stream) {
    PLNLP_REQUIRE(H > 0 && n_seg >= 0, PLNLP_E_SIZE);
    if (n_seg == 0) return 0;
    PLNLP_REQUIRE(seg_ptr && grad_h, PLNLP_E_NULL);
    PLNLP_REQUIRE(ldg >= H && (!g || ldg_in >= H), PLNLP_E_SIZE);
    const int vec = pick_vec(H, {ldg, g ? ldg_in : 4}, {g, grad_h});
    const unsigned grid = static_cast<unsigned>(ceil_div(n_seg, 8));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(V) \
    row_scatter_sorted_kernel<V><<<grid, 256, 0, st>>>(g, ldg_in, static_cast<int>(H), seg_ptr, n_seg, entry, grad_h, ldg)
    DISPATCH_VEC(vec, CALL);
#undef CALL
    PLNLP_LAUNCH_CHECK();
    return 0;
}
