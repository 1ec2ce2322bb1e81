// TMA-fed, persistent, warp-specialised tcgen05 GEMM for sm_100a:   C[M, N] = epi( A[M, K] . B^T ),  N <= 512.
//
// The shapes this kernel exists for are the TALL-SKINNY dense layers of the full-graph encoder
// (/root/reference/plnlp/layer.py:20,23 on the citation2-shape graph: 2.9 M x 200 x 178, 2.9 M x 50 x 200 -- M huge,
// N and K a couple of hundred).  They are bound by HBM (read A once, write C once) and, in error-compensated
// 3xTF32, equally by the tensor pipe; the register-path loader of gemm_tcgen05*.cu leaves them at ~21 % of the HBM
// roof because its short main loop (6 - 13 k-slabs) exposes prologue and epilogue.  Here:
//
//   * A tiles (128 rows x 32 fp32 of K = one 128-byte swizzle row each) arrive by TMA (cp.async.bulk.tensor,
//     SWIZZLE_128B, zero fill out of bounds) into an SA-deep ring; four CONVERT warps turn each raw tile into its
//     tf32 hi part in place (rn) and the lo part x - hi next to it (the 3xTF32 split; elementwise, so the swizzle
//     never has to be undone);
//   * B (the weight, N x K) is split ONCE per call by a small prologue kernel into K-major hi / lo copies in the
//     caller's workspace (whatever its layout: W for y = x W^T, W^T for dX = dY W) and both copies are TMA-loaded
//     per k-slab into a 2-deep ring of their own -- B comes out of L2, A out of HBM, so the rings are decoupled and
//     shared memory goes to A depth;
//   * one elected thread issues tcgen05.mma kind::tf32 (hi.hi + hi.lo + lo.hi) into a DOUBLE-BUFFERED TMEM
//     accumulator (2 x 256 columns): while four EPILOGUE warps drain tile t (tcgen05.ld -> bias / relu / dropout /
//     relu-grad mask / beta*C -> shared-memory turn-around -> coalesced 128-byte row segments to global), the main
//     loop of tile t + 1 is already running;
//   * CTAs are persistent (one per SM, tiles round-robin), so barriers, TMEM and descriptors are set up once.
//
// Numerics are those of gemm_tcgen05*.cu (hi = cvt.rna.tf32, lo = x - hi, same MMA order per k-step), K per
// accumulator is bounded by the caller (split-k shapes stay on the older kernels).
#include <cuda.h>

#include <cstdlib>

#include "gemm_tc_common.cuh"

namespace plnlp {

namespace {

using namespace tcgemm;

constexpr int TK = 32;                       // fp32 of K per slab: one 128-byte swizzle row
constexpr int A_TILE = TBM * 128;            // 16 KB
// warp 0: TMA producer of A, warp 1: TMA producer of B (independent rings: the deep A ring must be able to run ahead of
// the 2-deep B ring), warp 2: MMA issuer + TMEM owner, warps 3-6: convert, warps 7-14: epilogue (warp % 4 = TMEM lane
// quarter, two warps per quarter)
constexpr int TMA_THREADS = 480;
constexpr int EPI_WARPS = 8;
constexpr int MAX_N = 512;
constexpr int MAX_SA = 6;
constexpr int TN = 256;                      // columns per tile (UMMA N max, TMEM columns per accumulator)

struct TmaGemmParams {
    TcGemmParams g;
    int n_box;          // rows of the B box = min(256, roundup16(N))
    int sa;             // A ring depth
    int64_t m_tiles, n_tiles;
    int k_slabs;
};

__global__ void __launch_bounds__(256) split_b_kernel(const float* __restrict__ B, int64_t ldb, int transb, int64_t N,
                                                      int64_t K, int64_t kp, float* __restrict__ hi,
                                                      float* __restrict__ lo) {
    // element (n, k) of the K-major operand: B[n*ldb + k] (transb: B is [N, K]) or B[k*ldb + n] (B is [K, N])
    const int64_t total = N * kp;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t n = i / kp, k = i % kp;
        float v = 0.0f;
        if (k < K) v = transb ? __ldg(B + n * ldb + k) : __ldg(B + k * ldb + n);
        const float h = tc::to_tf32(v);
        hi[i] = h;
        lo[i] = v - h;
    }
}

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// Epilogue of one 128 x n_mma tile by the EIGHT epilogue warps: warp % 4 = TMEM lane quarter (32 rows), the two warps
// of a quarter take alternate 32-column chunks.  tcgen05.ld hands every thread 32 consecutive columns of ITS row;
// storing those straight to global memory makes every warp-level store touch 32 different 128-byte lines, and with one
// warp per quarter the chunk-after-chunk latency chain (tcgen05.ld -> bias -> turn-around -> stores: ~2.2 us per
// chunk measured) made the epilogue, not the main loop, the bound of the kernel (2.74 ms with it, 1.09 ms without).
// So: two warps per quarter, the bias vector in shared memory, and each warp turns its block around through 2 KB of
// shared memory, 16 columns at a time (16-byte chunks XOR-swizzled by row pair: the row-wise writes and the
// 8-rows-per-instruction reads are both bank-conflict free), writing 8 rows x 64 contiguous bytes per instruction.
//
// EPI selects how much epilogue code the instantiation carries: 0 = plain store, 1 = bias and / or relu (the encoder
// layers), 2 = everything (beta*C, dropout, relu-grad mask).  This is not cosmetic: with the generic epilogue (Philox
// rounds and all branches unrolled for 32 columns) the kernel was 9 000 SASS instructions, the five warp roles
// thrashed the instruction cache and the epilogue warps spent their time in "no instruction" stalls.
template <int EPI>
__device__ __forceinline__ void epilogue_tile_coalesced(const TcGemmParams& p, uint32_t tmem_d, int64_t m0, int64_t n0,
                                                        int n_mma, int q, int chunk0, int lane, uint32_t stage,
                                                        const float* bias_s) {
    const int64_t r = m0 + q * 32 + lane;
    const bool plain = p.beta == 0.0f && p.bias == nullptr && p.act == PLNLP_ACT_NONE;
    const float keep_scale = 1.0f / (1.0f - p.drop_p);
    const bool c_vec = (p.ldc % 4 == 0) && (reinterpret_cast<uintptr_t>(p.C) % 16 == 0);
    const bool vec_epi = (p.N % 4 == 0) && c_vec && (!p.aux || ((p.ldaux % 4 == 0) && reinterpret_cast<uintptr_t>(p.aux) % 16 == 0));
    const int sub = lane >> 2, ch = lane & 3;
    const uint32_t wr = stage + lane * 64;                       // this thread's row of the [32][16] staging block
    const int wsw = (lane >> 1) & 3;
    const uint32_t trow = tmem_d + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t nxt[32];
    if (chunk0 * 32 < n_mma) {
        tc::tmem_ld_32x32_nowait(trow + static_cast<uint32_t>(chunk0 * 32), nxt);
        tc::tmem_ld_wait();
    }
    for (int cb = chunk0 * 32; cb < n_mma; cb += 64) {
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(nxt[e]);
        // the accumulator columns of this warp's NEXT chunk are requested now and waited for at the end of the body
        if (cb + 64 < n_mma) tc::tmem_ld_32x32_nowait(trow + static_cast<uint32_t>(cb + 64), nxt);
        const int64_t c0 = n0 + cb;
        if constexpr (EPI == 1) {
            if (bias_s) {
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + e);
                    v[e] += b4.x; v[e + 1] += b4.y; v[e + 2] += b4.z; v[e + 3] += b4.w;
                }
            }
            if (p.act == PLNLP_ACT_RELU) {
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.0f);
            }
        } else if constexpr (EPI == 2) {
            if (r < p.M && !plain) tc_epi_apply32(p, r, c0, v, vec_epi, keep_scale, bias_s);
        }
        if (c_vec) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {                        // columns c0 + 16 h .. + 15
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    sts128(wr + ((j ^ wsw) << 4), v[16 * h + 4 * j], v[16 * h + 4 * j + 1], v[16 * h + 4 * j + 2],
                           v[16 * h + 4 * j + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = i * 8 + sub;
                    const float4 t = lds128(stage + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
                    const int64_t gr = m0 + q * 32 + rr, gc = c0 + 16 * h + ch * 4;
                    if (gr < p.M) {
                        float* dst = p.C + gr * p.ldc + gc;
                        if (gc + 3 < p.N) {
                            *reinterpret_cast<float4*>(dst) = t;
                        } else {
                            if (gc < p.N) dst[0] = t.x;
                            if (gc + 1 < p.N) dst[1] = t.y;
                            if (gc + 2 < p.N) dst[2] = t.z;
                        }
                    }
                }
                __syncwarp();
            }
        } else {
            // C rows are not 16-byte aligned (e.g. a [M, 50] matrix): the same turn-around with 4-byte accesses, every
            // instruction writing 2 rows x 64 contiguous bytes
            const int sub2 = lane >> 4, col = lane & 15;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    sts128(wr + ((j ^ wsw) << 4), v[16 * h + 4 * j], v[16 * h + 4 * j + 1], v[16 * h + 4 * j + 2],
                           v[16 * h + 4 * j + 3]);
                __syncwarp();
                const int64_t gc = c0 + 16 * h + col;
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int rr = i * 2 + sub2;
                    float t;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t)
                                 : "r"(stage + rr * 64 + (((col >> 2) ^ ((rr >> 1) & 3)) << 4) + ((col & 3) << 2)) : "memory");
                    const int64_t gr = m0 + q * 32 + rr;
                    if (gr < p.M && gc < p.N) p.C[gr * p.ldc + gc] = t;
                }
                __syncwarp();
            }
        }
        if (cb + 64 < n_mma) tc::tmem_ld_wait();
    }
}

template <int EPI>
__global__ void __launch_bounds__(TMA_THREADS, 1)
    gemm_tma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_bh,
                    const __grid_constant__ CUtensorMap tm_bl, const TmaGemmParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_a[MAX_SA], conv_a[MAX_SA], empty_a[MAX_SA];
    __shared__ __align__(8) uint64_t full_b[2], empty_b[2], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_holder;
    __shared__ __align__(16) float bias_s[MAX_N];

    const TcGemmParams& p = P.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int SA = P.sa;
    const uint32_t b_bytes = static_cast<uint32_t>(P.n_box) * 128u;      // one B part of one stage
    // 1024-byte aligned carve-up: [A hi | A lo] x SA, then [B hi | B lo] x 2
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_base = smem;
    uint8_t* b_base = smem + static_cast<size_t>(SA) * 2 * A_TILE;
    uint8_t* stage_base = b_base + 4 * static_cast<size_t>(P.n_box) * 128;     // 8 x 2 KB epilogue staging
    auto a_hi = [&](int s) { return a_base + static_cast<size_t>(s) * 2 * A_TILE; };
    auto a_lo = [&](int s) { return a_base + static_cast<size_t>(s) * 2 * A_TILE + A_TILE; };
    auto b_hi = [&](int s) { return b_base + static_cast<size_t>(s) * 2 * b_bytes; };
    auto b_lo = [&](int s) { return b_base + static_cast<size_t>(s) * 2 * b_bytes + b_bytes; };

    if (tid == 0) {
        for (int s = 0; s < SA; ++s) {
            tc::mbar_init(&full_a[s], 1);
            tc::mbar_init(&conv_a[s], 128);
            tc::mbar_init(&empty_a[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&full_b[s], 1);
            tc::mbar_init(&empty_b[s], 1);
            tc::mbar_init(&tmem_full[s], 1);
            tc::mbar_init(&tmem_empty[s], EPI_WARPS * 32);
        }
        tc::mbar_fence_init();
    }
    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tm_a);
        tc::tma_prefetch_desc(&tm_bh);
        tc::tma_prefetch_desc(&tm_bl);
    }
    if (warp == 2) tc::tmem_alloc<512>(&tmem_holder);
    if (p.bias)
        for (int c = tid; c < MAX_N; c += TMA_THREADS) bias_s[c] = c < p.N ? __ldg(p.bias + c) : 0.0f;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_holder;

    const int64_t total_tiles = P.m_tiles * P.n_tiles;
    const int KS = P.k_slabs;
    const bool split = p.passes == 3;

    if (warp == 0) {
        // ============================ TMA producer: A ring ============================
        if (lane == 0) {
            uint32_t ia = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m0 = static_cast<int>((tile / P.n_tiles) * TBM);
                for (int ks = 0; ks < KS; ++ks, ++ia) {
                    const int sa = ia % SA;
                    tc::mbar_wait(&empty_a[sa], ((ia / SA) & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(&full_a[sa], A_TILE);
                    tc::tma_load_2d(a_hi(sa), &tm_a, ks * TK, m0, &full_a[sa]);
                }
            }
        }
    } else if (warp == 1) {
        // ============================ TMA producer: B ring (hi and lo copies of the weight) ============
        if (lane == 0) {
            uint32_t ib = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n0 = static_cast<int>((tile % P.n_tiles) * TN);
                for (int ks = 0; ks < KS; ++ks, ++ib) {
                    const int sb = ib & 1;
                    tc::mbar_wait(&empty_b[sb], ((ib >> 1) & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(&full_b[sb], split ? 2 * b_bytes : b_bytes);
                    tc::tma_load_2d(b_hi(sb), &tm_bh, ks * TK, n0, &full_b[sb]);
                    if (split) tc::tma_load_2d(b_lo(sb), &tm_bl, ks * TK, n0, &full_b[sb]);
                }
            }
        }
    } else if (warp == 2) {
        // ============================ MMA issuer ============================
        uint32_t ia = 0, ib = 0, tl = 0;
        for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
            const int64_t n0 = (tile % P.n_tiles) * TN;
            const int64_t n_rem = ((p.N - n0 + 15) / 16) * 16;
            const int n_mma = n_rem < TN ? static_cast<int>(n_rem) : TN;
            const uint32_t idesc = tc::make_idesc_tf32(TBM, n_mma, 0, 0);
            const uint32_t acc = tl & 1;
            const uint32_t tmem_d = tmem_base + acc * TN;
            tc::mbar_wait(&tmem_empty[acc], ((tl >> 1) & 1) ^ 1);        // the epilogue has drained this accumulator
            tc::fence_after_sync();
            for (int ks = 0; ks < KS; ++ks, ++ia, ++ib) {
                const int sa = ia % SA, sb = ib & 1;
                tc::mbar_wait(&conv_a[sa], (ia / SA) & 1);               // hi / lo of the A tile are in place
                tc::mbar_wait(&full_b[sb], (ib >> 1) & 1);
                tc::fence_after_sync();
                if (lane == 0) {
                    const uint32_t ah = tc::smem_u32(a_hi(sa)), al = tc::smem_u32(a_lo(sa));
                    const uint32_t bh = tc::smem_u32(b_hi(sb)), bl = tc::smem_u32(b_lo(sb));
#pragma unroll
                    for (int j = 0; j < TK / 8; ++j) {
                        const uint64_t dah = tc::make_smem_desc_sw128(ah + j * 32);
                        const uint64_t dbh = tc::make_smem_desc_sw128(bh + j * 32);
                        tc::mma_tf32_ss(tmem_d, dah, dbh, idesc, (ks | j) != 0);
                        if (split) {
                            const uint64_t dal = tc::make_smem_desc_sw128(al + j * 32);
                            const uint64_t dbl = tc::make_smem_desc_sw128(bl + j * 32);
                            tc::mma_tf32_ss(tmem_d, dah, dbl, idesc, 1u);
                            tc::mma_tf32_ss(tmem_d, dal, dbh, idesc, 1u);
                        }
                    }
                    tc::mma_commit(&empty_a[sa]);
                    tc::mma_commit(&empty_b[sb]);
                    if (ks == KS - 1) tc::mma_commit(&tmem_full[acc]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 7) {
        // ============================ convert warps: raw fp32 tile -> tf32 hi (in place) + lo ============
        const int t = tid - 96;                                          // 0..127
        uint32_t ia = 0;
        for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            for (int ks = 0; ks < KS; ++ks, ++ia) {
                const int sa = ia % SA;
                tc::mbar_wait(&full_a[sa], (ia / SA) & 1);
                float4* hi = reinterpret_cast<float4*>(a_hi(sa));
                float4* lo = reinterpret_cast<float4*>(a_lo(sa));
#pragma unroll
                for (int i = 0; i < A_TILE / 16 / 128; ++i) {            // 8 chunks of 16 bytes per thread
                    const int c = t + 128 * i;
                    const float4 v = hi[c];
                    float4 h, l;
                    h.x = tc::to_tf32(v.x); h.y = tc::to_tf32(v.y); h.z = tc::to_tf32(v.z); h.w = tc::to_tf32(v.w);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                    hi[c] = h;
                    if (split) lo[c] = l;
                }
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(&conv_a[sa]);
            }
        }
    } else {
        // ============================ epilogue warps ============================
        uint32_t tl = 0;
        for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
            const int64_t m0 = (tile / P.n_tiles) * TBM, n0 = (tile % P.n_tiles) * TN;
            const int64_t n_rem = ((p.N - n0 + 15) / 16) * 16;
            const int n_mma = n_rem < TN ? static_cast<int>(n_rem) : TN;
            const uint32_t acc = tl & 1;
            tc::mbar_wait(&tmem_full[acc], (tl >> 1) & 1);
            tc::fence_after_sync();
            epilogue_tile_coalesced<EPI>(p, tmem_base + acc * TN, m0, n0, n_mma, warp & 3, (warp - 7) >> 2, lane,
                                         tc::smem_u32(stage_base) + (warp - 7) * 2048, p.bias ? bias_s : nullptr);
            tc::fence_before_sync();
            tc::mbar_arrive(&tmem_empty[acc]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc<512>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// [rows, cols] fp32 row-major matrix with leading dimension ld (elements); box = box_rows x 32 columns
int make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return PLNLP_E_UNSUPPORTED;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
    const cuuint32_t box[2] = {TK, static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : PLNLP_E_UNSUPPORTED;
}

}  // namespace
}  // namespace plnlp

// bytes of caller-owned workspace for the hi / lo copies of the B operand
extern "C" int64_t plnlp_gemm_tf32_tma_workspace_bytes(int64_t N, int64_t K) {
    const int64_t kp = (K + 3) / 4 * 4;
    return 2 * N * kp * 4 + 256;
}

// C = act(A @ op(B) + beta*C + bias) with A [M, K] row-major (K contiguous), op(B) = B^T for transb = 1 (B is [N, K])
// or B for transb = 0 (B is [K, N]); same epilogue contract as plnlp_gemm_tf32_2cta.  Requirements (else
// PLNLP_E_UNSUPPORTED and the caller uses another kernel): lda % 4 == 0, A 16-byte aligned, 32 <= K, N <= 512.
extern "C" int plnlp_gemm_tf32_tma(int passes, int transb, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                                   const float* B, int64_t ldb, float* C, int64_t ldc, float beta, const float* bias,
                                   int act, const float* aux, int64_t ldaux, float drop_p, uint64_t seed,
                                   float* workspace, int64_t workspace_bytes, void* stream) {
    using namespace plnlp;
    using namespace plnlp::tcgemm;
    PLNLP_REQUIRE(passes == 1 || passes == 3, PLNLP_E_UNSUPPORTED);
    PLNLP_REQUIRE(M >= 0 && N >= 0 && K >= 0, PLNLP_E_SIZE);
    if (M == 0 || N == 0) return 0;
    PLNLP_REQUIRE(A && B && C && workspace, PLNLP_E_NULL);
    PLNLP_REQUIRE(lda >= K && ldb >= (transb ? K : N) && ldc >= N, PLNLP_E_SIZE);
    PLNLP_REQUIRE(act >= 0 && act <= 2 && drop_p >= 0.0f && drop_p < 1.0f, PLNLP_E_SIZE);
    if (act == PLNLP_ACT_RELU_GRAD) PLNLP_REQUIRE(aux && ldaux >= N, PLNLP_E_NULL);
    PLNLP_REQUIRE(K >= TK && N <= 2 * TN && M < (int64_t(1) << 31) - TBM, PLNLP_E_UNSUPPORTED);
    PLNLP_REQUIRE((lda % 4 == 0) && aligned(A, 16), PLNLP_E_UNSUPPORTED);
    PLNLP_REQUIRE(workspace_bytes >= plnlp_gemm_tf32_tma_workspace_bytes(N, K), PLNLP_E_WORKSPACE);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    const int64_t kp = (K + 3) / 4 * 4;
    float* bhi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    float* blo = bhi + N * kp;
    {
        const int64_t total = N * kp;
        const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div(total, 256), 1184));
        split_b_kernel<<<blocks, 256, 0, st>>>(B, ldb, transb, N, K, kp, bhi, blo);
        PLNLP_LAUNCH_CHECK();
    }

    TmaGemmParams P{};
    TcGemmParams& p = P.g;
    p.M = M; p.N = N; p.K = K; p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
    p.beta = beta; p.bias = bias; p.act = act; p.aux = aux; p.ldaux = ldaux; p.drop_p = drop_p; p.seed = seed;
    p.passes = passes; p.split_k = 1; p.k_per_split = K;
    const int64_t n16 = ceil_div(N, 16) * 16;
    P.n_box = static_cast<int>(std::min<int64_t>(n16, TN));
    P.m_tiles = ceil_div(M, TBM);
    P.n_tiles = ceil_div(N, TN);
    P.k_slabs = static_cast<int>(ceil_div(K, TK));
    const int b_ring = 2 * 2 * P.n_box * 128;                       // 2 stages x (hi + lo)
    constexpr int EPI_STAGE = EPI_WARPS * 2048;                     // one 32 x 16 fp32 block per epilogue warp
    const int budget = 227 * 1024 - 1024 /* alignment slack */ - 3072 /* static */ - b_ring - EPI_STAGE;
    P.sa = std::max(2, std::min(MAX_SA, budget / (2 * A_TILE)));
    const int smem_bytes = P.sa * 2 * A_TILE + b_ring + EPI_STAGE + 1024;

    CUtensorMap tm_a, tm_bh, tm_bl;
    int rc = make_map(&tm_a, A, M, K, lda, TBM);
    if (rc == 0) rc = make_map(&tm_bh, bhi, N, K, kp, P.n_box);
    if (rc == 0) rc = make_map(&tm_bl, blo, N, K, kp, P.n_box);
    if (rc != 0) return rc;

    static bool configured = false;
    if (!configured) {
        const void* kerns[3] = {reinterpret_cast<const void*>(&gemm_tma_kernel<0>), reinterpret_cast<const void*>(&gemm_tma_kernel<1>),
                                reinterpret_cast<const void*>(&gemm_tma_kernel<2>)};
        for (const void* kern : kerns) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 3072);
            if (e != cudaSuccess) return static_cast<int>(e);
        }
        configured = true;
    }
    static const int sa_env = [] { const char* e = getenv("PLNLP_TMA_SA"); return e ? atoi(e) : 0; }();
    if (sa_env >= 2 && sa_env <= P.sa) P.sa = sa_env;
    const int64_t total_tiles = P.m_tiles * P.n_tiles;
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(total_tiles, kNumSM));
    const bool plain = beta == 0.0f && bias == nullptr && act == PLNLP_ACT_NONE;
    const bool light = beta == 0.0f && drop_p == 0.0f && (act == PLNLP_ACT_NONE || act == PLNLP_ACT_RELU);
    if (plain) gemm_tma_kernel<0><<<grid, TMA_THREADS, smem_bytes, st>>>(tm_a, tm_bh, tm_bl, P);
    else if (light) gemm_tma_kernel<1><<<grid, TMA_THREADS, smem_bytes, st>>>(tm_a, tm_bh, tm_bl, P);
    else gemm_tma_kernel<2><<<grid, TMA_THREADS, smem_bytes, st>>>(tm_a, tm_bh, tm_bl, P);
    PLNLP_LAUNCH_CHECK();
    return 0;
}
