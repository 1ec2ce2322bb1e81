// TMA-fed, persistent, warp-specialised tcgen05 GEMM for WEIGHT GRADIENTS over a huge row count:
//
//     C[M, N] = A^T . B      A [K, M], B [K, N] row-major (both "MN-major": K is the slow index), M, N <= 256, K huge
//
// This is dW = dY^T [A_hat x | 1] of the first encoder layer on the citation2-shape graph
// (/root/reference/plnlp/layer.py:20,23 backward; 200 x 179 x 2 927 963): 4.4 GB streamed once, 0.21 TFLOP.  On the
// register-path CTA-pair kernel (gemm_tcgen05_2cta.cu) it ran 2.5 - 2.7 ms = 12 - 15 % of the training step: both operands
// need the 4 x 4 register transposes of the MN-major loader, the 256 x 256 pair tile is 55 % padding, and the
// round-toward-zero accumulation of the tensor core caps K per accumulator at 1088, i.e. 2692 split-k CTAs each with
// their own prologue, epilogue and a 143 KB partial.
//
// Here nothing is transposed: tcgen05.mma takes MN-major operands directly (instruction-descriptor bits 15 / 16), and
// that is exactly what a TMA box of 32 fp32 along M (or N) x 16 rows of K lands in shared memory with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B -- the one MN-major layout tcgen05 takes for 32-bit operands: rows of 128 bytes,
// 32-byte chunks XOR-ed with (row % 4); the atoms of the next 32 M-elements are LBO bytes further, the next 4 K-rows
// SBO = 512 bytes.
//   * warp 0: TMA producer -- per 16-row k-slab one box per 32-column atom of A and of B into a 4-deep ring;
//   * warps 2-5: convert -- raw fp32 -> tf32 hi (in place) + lo = x - hi next to it (3xTF32, elementwise, so the
//     swizzle never has to be undone);
//   * warp 1: one elected thread issues, per 8 K-rows and per 128-row half of M, hi.hi + hi.lo + lo.hi into that
//     half's TMEM accumulator (2 x 256 columns);
//   * warps 6-13: epilogue -- every UNIT of 1088 K-rows (the RZ error budget of one accumulator) the two accumulators
//     are added (RN, CUDA cores) into the CTA's running partial in global memory (L2 resident);
//   * CTAs are persistent, one per SM, each owns a contiguous range of units; a small second kernel adds the (at most
//     148) CTA partials in CTA order.  Deterministic.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "gemm_tc_common.cuh"

namespace plnlp {

namespace {

using namespace tcgemm;

constexpr int TN_KB = 16;                        // K-rows per slab (two swizzle atoms of 8 rows)
constexpr int TN_ATOM = TN_KB * 128;             // bytes of one 32-column atom of one slab: 2 KB
constexpr int TN_UNIT = 1088;                    // K-rows per accumulator (68 slabs), see _ops.TF32X3_KCAP
constexpr int TN_STAGES = 4;
constexpr int TN_THREADS = 448;                  // 14 warps: TMA, MMA, 4 convert, 8 epilogue
constexpr int TN_CONV = 128;                     // convert threads
constexpr int TN_EPI = 256;                      // epilogue threads
constexpr int TN_PLD = 256;                      // leading dimension of a CTA partial [256][256]

struct TnParams {
    int64_t M, N, K;
    float* C; int64_t ldc;
    float* partial;          // [gridDim.x][256][256]
    int ma, ma_live, na;     // 32-column atoms of A in the stage layout (4 or 8) / that exist (ceil(M / 32)); atoms of B
    int a_full, b_full;      // atoms that lie wholly inside the matrix: fetched by ONE 3-D box per operand and slab
    int n_mma;               // roundup16(N)
    int64_t n_units;
    int passes;
    int stages;              // ring depth (2 .. TN_STAGES): as many as fit the shared memory
    int dbg;                 // experiment switches (PLNLP_TN_DEBUG): bit 0/1 = clear the MN-major bit of A / B, bit 2 = swap LBO and SBO, 8 = plain SWIZZLE_128B, 16 = no MMAs, 32 = no convert (timing decomposition)
};

// MN-major tf32 operand.  The only shared-memory layout tcgen05 accepts for MN-major 32-bit operands is "128-byte
// swizzle with a 32-byte base" (layout type 1; plain SWIZZLE_128B makes the MMA return zeros): rows of 128 bytes
// (32 MN-elements), the four 32-byte chunks of a row XOR-ed with (row % 4), i.e. swizzle atoms of 4 K-rows = 512 bytes
// -- exactly what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  LBO = bytes between the atoms of the next 32
// MN-elements, SBO = bytes between consecutive 4-row K groups.
__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                            uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;             // descriptor version for sm_100
    d |= static_cast<uint64_t>(layout_type) << 61;
    return d;
}

__global__ void __launch_bounds__(TN_THREADS, 1)
    gemm_tma_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                       const __grid_constant__ CUtensorMap tm_a3, const __grid_constant__ CUtensorMap tm_b3,
                       const TnParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TN_STAGES], conv_bar[TN_STAGES], empty_bar[TN_STAGES], tmem_full, tmem_empty;
    __shared__ uint32_t tmem_holder;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int atoms = P.ma + P.na;
    const uint32_t part_bytes = static_cast<uint32_t>(atoms) * TN_ATOM;            // hi (or lo) part of one stage
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    auto hi_part = [&](int s) { return smem + static_cast<size_t>(s) * 2 * part_bytes; };
    auto lo_part = [&](int s) { return smem + static_cast<size_t>(s) * 2 * part_bytes + part_bytes; };

    if (tid == 0) {
        for (int s = 0; s < TN_STAGES; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&conv_bar[s], TN_CONV);
            tc::mbar_init(&empty_bar[s], 1);
        }
        tc::mbar_init(&tmem_full, 1);
        tc::mbar_init(&tmem_empty, TN_EPI);
        tc::mbar_fence_init();
    }
    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tm_a);
        tc::tma_prefetch_desc(&tm_b);
        tc::tma_prefetch_desc(&tm_a3);
        tc::tma_prefetch_desc(&tm_b3);
    }
    if (warp == 1) tc::tmem_alloc<512>(&tmem_holder);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_holder;

    // this CTA's contiguous range of units
    const int64_t u0 = P.n_units * blockIdx.x / gridDim.x, u1 = P.n_units * (blockIdx.x + 1) / gridDim.x;
    const bool split = P.passes == 3;

    if (warp == 0) {
        // ============================ TMA producer ============================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t u = u0; u < u1; ++u) {
                const int64_t k_beg = u * TN_UNIT, k_end = min(P.K, k_beg + TN_UNIT);
                for (int64_t k0 = k_beg; k0 < k_end; k0 += TN_KB, ++it) {
                    const int s = it % P.stages;
                    tc::mbar_wait(&empty_bar[s], ((it / P.stages) & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(P.ma_live + P.na) * TN_ATOM);
                    uint8_t* dst = hi_part(s);
                    // the producer thread is bound by the number of TMA instructions it issues (13 boxes of 2 KB per
                    // slab: 1.33 ms for the ring alone): the atoms that lie wholly inside the matrix come as ONE 3-D box
                    // (32 columns x 16 rows x atoms), only a ragged last atom needs its own zero-filled 2-D box
                    if (P.a_full > 0) tc::tma_load_3d(dst, &tm_a3, 0, static_cast<int>(k0), 0, &full_bar[s]);
                    for (int a = P.a_full; a < P.ma_live; ++a)
                        tc::tma_load_2d(dst + a * TN_ATOM, &tm_a, a * 32, static_cast<int>(k0), &full_bar[s]);
                    if (P.b_full > 0)
                        tc::tma_load_3d(dst + P.ma * TN_ATOM, &tm_b3, 0, static_cast<int>(k0), 0, &full_bar[s]);
                    for (int b = P.b_full; b < P.na; ++b)
                        tc::tma_load_2d(dst + (P.ma + b) * TN_ATOM, &tm_b, b * 32, static_cast<int>(k0), &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        const uint32_t idesc = tc::make_idesc_tf32(TBM, P.n_mma, (P.dbg & 1) ? 0 : 1, (P.dbg & 2) ? 0 : 1);
        // PLNLP_TN_DEBUG bit 3: the plain SWIZZLE_128B experiment (returns zeros)
        const uint32_t LT = (P.dbg & 8) ? 2u : 1u;
        const uint32_t sbo0 = (P.dbg & 8) ? 1024u : 512u;
        const uint32_t LBO = (P.dbg & 4) ? sbo0 : static_cast<uint32_t>(TN_ATOM), SBO = (P.dbg & 4) ? static_cast<uint32_t>(TN_ATOM) : sbo0;
        const int halves = P.ma > 4 ? 2 : 1;
        uint32_t it = 0, ul = 0;
        for (int64_t u = u0; u < u1; ++u, ++ul) {
            const int64_t k_beg = u * TN_UNIT, k_end = min(P.K, k_beg + TN_UNIT);
            tc::mbar_wait(&tmem_empty, (ul & 1) ^ 1);                   // the epilogue has drained the accumulators
            tc::fence_after_sync();
            bool first = true;
            for (int64_t k0 = k_beg; k0 < k_end; k0 += TN_KB, ++it) {
                const int s = it % P.stages;
                tc::mbar_wait(&conv_bar[s], (it / P.stages) & 1);      // hi / lo of the slab are in place
                tc::fence_after_sync();
                if (lane == 0) {
                    const uint32_t hi = tc::smem_u32(hi_part(s)), lo = tc::smem_u32(lo_part(s));
                    const uint32_t b_off = static_cast<uint32_t>(P.ma) * TN_ATOM;
#pragma unroll
                    for (int kg = 0; kg < ((P.dbg & 16) ? 0 : TN_KB / 8); ++kg) {
                        const uint64_t dbh = make_smem_desc_sw128_mn(hi + b_off + kg * 1024, LBO, SBO, LT);
                        const uint64_t dbl = make_smem_desc_sw128_mn(lo + b_off + kg * 1024, LBO, SBO, LT);
                        for (int h = 0; h < halves; ++h) {
                            const uint32_t a_off = static_cast<uint32_t>(h) * 4 * TN_ATOM + kg * 1024;
                            const uint64_t dah = make_smem_desc_sw128_mn(hi + a_off, LBO, SBO, LT);
                            const uint32_t d = tmem_base + static_cast<uint32_t>(h) * 256;
                            tc::mma_tf32_ss(d, dah, dbh, idesc, first ? 0u : 1u);
                            if (split) {
                                const uint64_t dal = make_smem_desc_sw128_mn(lo + a_off, LBO, SBO, LT);
                                tc::mma_tf32_ss(d, dah, dbl, idesc, 1u);
                                tc::mma_tf32_ss(d, dal, dbh, idesc, 1u);
                            }
                        }
                        first = false;
                    }
                    tc::mma_commit(&empty_bar[s]);
                    if (k0 + TN_KB >= k_end) tc::mma_commit(&tmem_full);
                }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // ============================ convert warps: raw fp32 -> tf32 hi (in place) + lo ============
        const int t = tid - 64;                                          // 0..127
        const int chunks = static_cast<int>(part_bytes / 16);           // 16-byte chunks of one part
        uint32_t it = 0;
        for (int64_t u = u0; u < u1; ++u) {
            const int64_t k_beg = u * TN_UNIT, k_end = min(P.K, k_beg + TN_UNIT);
            for (int64_t k0 = k_beg; k0 < k_end; k0 += TN_KB, ++it) {
                const int s = it % P.stages;
                tc::mbar_wait(&full_bar[s], (it / P.stages) & 1);
                float4* hi = reinterpret_cast<float4*>(hi_part(s));
                float4* lo = reinterpret_cast<float4*>(lo_part(s));
#pragma unroll 4
                for (int c = t; c < ((P.dbg & 32) ? 0 : chunks); c += TN_CONV) {
                    const float4 v = hi[c];
                    float4 h, l;
                    h.x = tc::to_tf32(v.x); h.y = tc::to_tf32(v.y); h.z = tc::to_tf32(v.z); h.w = tc::to_tf32(v.w);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                    hi[c] = h;
                    if (split) lo[c] = l;
                }
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(&conv_bar[s]);
            }
        }
    } else {
        // ============================ epilogue warps: accumulators -> the CTA's running partial ============
        // a warp can only read the TMEM lane quarter (warp id % 4): warps 6..9 take the quarters 2, 3, 0, 1 of the first
        // 128-row half of M, warps 10..13 those of the second
        const int q = warp & 3, h = (warp - 6) >> 2;
        const int halves = P.ma > 4 ? 2 : 1;
        float* prow = P.partial + (static_cast<int64_t>(blockIdx.x) * 256 + h * 128 + q * 32 + lane) * TN_PLD;
        uint32_t ul = 0;
        for (int64_t u = u0; u < u1; ++u, ++ul) {
            tc::mbar_wait(&tmem_full, ul & 1);
            tc::fence_after_sync();
            if (h < halves) {
                const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(h) * 256;
                for (int cb = 0; cb < P.n_mma; cb += 32) {
                    float v[32];
                    tc::tmem_ld_32x32(trow + static_cast<uint32_t>(cb), v);
                    float4* dst = reinterpret_cast<float4*>(prow + cb);
                    if (u != u0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 o = dst[i];
                            v[4 * i] = __fadd_rn(o.x, v[4 * i]); v[4 * i + 1] = __fadd_rn(o.y, v[4 * i + 1]);
                            v[4 * i + 2] = __fadd_rn(o.z, v[4 * i + 2]); v[4 * i + 3] = __fadd_rn(o.w, v[4 * i + 3]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
            tc::fence_before_sync();
            tc::mbar_arrive(&tmem_empty);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<512>(tmem_base);
}

// C[m, n] = sum over the CTA partials, in CTA order
__global__ void __launch_bounds__(256) gemm_tma_tn_reduce_kernel(const float* __restrict__ partial, int n_part, int64_t M,
                                                                 int64_t N, float* __restrict__ C, int64_t ldc) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= M * N) return;
    const int64_t m = i / N, n = i % N;
    float acc = 0.0f;
    for (int g = 0; g < n_part; ++g) acc = __fadd_rn(acc, partial[(static_cast<int64_t>(g) * 256 + m) * TN_PLD + n]);
    C[m * ldc + n] = acc;
}

typedef CUresult (*EncodeTiledFnTn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFnTn encode_fn_tn() {
    static EncodeTiledFnTn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFnTn>(f);
    }();
    return fn;
}

// [rows = K, cols] fp32 row-major matrix with leading dimension ld (elements); box = 32 columns x TN_KB rows
int make_map_tn(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, bool plain_sw128) {
    EncodeTiledFnTn fn = encode_fn_tn();
    if (!fn) return PLNLP_E_UNSUPPORTED;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
    const cuuint32_t box[2] = {32, TN_KB};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          plain_sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : PLNLP_E_UNSUPPORTED;
}

// the `atoms` 32-column atoms that lie wholly inside a [K, cols] matrix as ONE box: dims (32 columns, K rows, atoms) with
// strides (ld * 4, 128) bytes -- the box lands atom-major (16 rows of 128 bytes per atom), the stage layout
int make_map_tn3(CUtensorMap* m, const float* base, int64_t rows, int atoms, int64_t ld, bool plain_sw128) {
    EncodeTiledFnTn fn = encode_fn_tn();
    if (!fn) return PLNLP_E_UNSUPPORTED;
    const cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(atoms)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 4, 128};
    const cuuint32_t box[3] = {32, TN_KB, static_cast<cuuint32_t>(atoms)};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          plain_sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : PLNLP_E_UNSUPPORTED;
}

}  // namespace
}  // namespace plnlp

// bytes of caller-owned workspace: one [256][256] fp32 partial per CTA (one CTA per SM)
extern "C" int64_t plnlp_gemm_tf32_tma_tn_workspace_bytes(void) {
    return static_cast<int64_t>(plnlp::kNumSM) * 256 * plnlp::TN_PLD * 4 + 256;
}

// C[M, N] = A^T @ B with A [K, M] (lda), B [K, N] (ldb) row-major; M, N <= 256; lda, ldb multiples of 4 and 16-byte
// aligned bases (TMA); passes = 3: error-compensated 3xTF32 (K per accumulator 1088, RN adds in between), 1: plain TF32.
extern "C" int plnlp_gemm_tf32_tma_tn(int passes, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                                      const float* B, int64_t ldb, float* C, int64_t ldc, float* workspace,
                                      int64_t workspace_bytes, void* stream) {
    using namespace plnlp;
    using namespace plnlp::tcgemm;
    PLNLP_REQUIRE(passes == 1 || passes == 3, PLNLP_E_UNSUPPORTED);
    PLNLP_REQUIRE(M > 0 && N > 0 && K > 0, PLNLP_E_SIZE);
    PLNLP_REQUIRE(A && B && C && workspace, PLNLP_E_NULL);
    PLNLP_REQUIRE(lda >= M && ldb >= N && ldc >= N, PLNLP_E_SIZE);
    PLNLP_REQUIRE(M <= 256 && N <= 256 && N >= 8 && K < (int64_t(1) << 31) - TN_UNIT, PLNLP_E_UNSUPPORTED);
    PLNLP_REQUIRE((lda % 4 == 0) && (ldb % 4 == 0) && aligned(A, 16) && aligned(B, 16), PLNLP_E_UNSUPPORTED);
    PLNLP_REQUIRE(workspace_bytes >= plnlp_gemm_tf32_tma_tn_workspace_bytes(), PLNLP_E_WORKSPACE);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    TnParams P{};
    P.M = M; P.N = N; P.K = K; P.C = C; P.ldc = ldc; P.passes = passes;
    P.partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    // Stage layout: [4 or 8 atoms of A | na atoms of B].  The MMA of 128-row half h reads A atoms 4h .. 4h + 3; atoms
    // past M are never loaded -- the accumulator rows they feed (rows of D depend only on the same rows of A) are
    // never read back, so whatever the shared memory holds there is harmless.
    P.ma_live = static_cast<int>(ceil_div(M, 32));
    P.ma = M > 128 ? 8 : 4;
    P.n_mma = static_cast<int>(ceil_div(N, 16) * 16);
    P.na = static_cast<int>(ceil_div(P.n_mma, 32));
    P.n_units = ceil_div(K, TN_UNIT);
    { const char* e = getenv("PLNLP_TN_DEBUG"); P.dbg = e ? atoi(e) : 0; }
    const int atoms_layout = P.ma + P.na;

    CUtensorMap tm_a, tm_b, tm_a3, tm_b3;
    int rc = make_map_tn(&tm_a, A, K, M, lda, (P.dbg & 8) != 0);
    if (rc == 0) rc = make_map_tn(&tm_b, B, K, N, ldb, (P.dbg & 8) != 0);
    if (rc != 0) return rc;
    // one 3-D box for the atoms wholly inside the matrix (PLNLP_TN_DEBUG bit 6: one 2-D box per atom, the first version)
    P.a_full = (P.dbg & 64) ? 0 : static_cast<int>(M / 32);
    P.b_full = (P.dbg & 64) ? 0 : static_cast<int>(N / 32);
    tm_a3 = tm_a;
    tm_b3 = tm_b;
    if (P.a_full > 0 && make_map_tn3(&tm_a3, A, K, P.a_full, lda, (P.dbg & 8) != 0) != 0) P.a_full = 0;
    if (P.b_full > 0 && make_map_tn3(&tm_b3, B, K, P.b_full, ldb, (P.dbg & 8) != 0) != 0) P.b_full = 0;

    P.stages = std::min(TN_STAGES, (227 * 1024 - 2048) / (2 * atoms_layout * TN_ATOM));
    PLNLP_REQUIRE(P.stages >= 2, PLNLP_E_UNSUPPORTED);
    const int smem_bytes = P.stages * 2 * atoms_layout * TN_ATOM + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tma_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024 - 1024);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured = true;
    }
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(P.n_units, kNumSM));
    gemm_tma_tn_kernel<<<grid, TN_THREADS, smem_bytes, st>>>(tm_a, tm_b, tm_a3, tm_b3, P);
    PLNLP_LAUNCH_CHECK();
    const int64_t total = M * N;
    gemm_tma_tn_reduce_kernel<<<static_cast<unsigned>(ceil_div(total, 256)), 256, 0, st>>>(P.partial, static_cast<int>(grid),
                                                                                         M, N, C, ldc);
    PLNLP_LAUNCH_CHECK();
    return 0;
}
