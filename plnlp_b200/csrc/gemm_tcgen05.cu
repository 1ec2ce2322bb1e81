// Tensor-core GEMM for sm_100a: tcgen05.mma kind::tf32, operands staged in shared memory in the
// canonical no-swizzle core-matrix layout, fp32 accumulator in TMEM, fused epilogue.
//
// Replaces torch.nn.Linear -> cuBLAS sgemm at /root/reference/plnlp/layer.py:20,23,82-86 and its two
// backward GEMMs on the tensor pipe while keeping fp32 parity:
//
//   passes = 3 ("3xTF32", error-compensated):  A = Ah + Al, B = Bh + Bl with Ah = rn_tf32(A),
//            Al = A - Ah (exact in fp32);  A.B ~= Ah.Bh + Ah.Bl + Al.Bh, all accumulated in fp32 in
//            TMEM.  The dropped Al.Bl term is 2^-22 relative.  The tensor core accumulates with
//            round-toward-zero, so the error grows linearly with the number of accumulation steps
//            (K/8): callers keep K per accumulator <= 1024 (split-k, partials reduced with RN adds on
//            the CUDA cores) which keeps the result within ~4e-6 relative -- inside the 1e-5 bar.
//   passes = 1: plain TF32 (~1e-3 relative), the stated fast path.
//
// CTA = 128 x BN output tile, 9 warps: warps 0-7 load A/B slabs from global memory (any of the four
// transpose combinations), split them into hi/lo parts and write the UMMA layout; warp 8 allocates
// TMEM and its elected lane issues the MMAs; full/empty mbarriers form a STAGES-deep ring
// (loaders -> MMA via fence.proxy.async + arrive, MMA -> loaders via tcgen05.commit).  Global loads
// run two slabs ahead of the shared-memory stores (two register buffers).  The shared-memory footprint
// is held to ~165 KB on purpose: the global loads travel through L1, and with a full 227 KB carve-out
// the ~7 KB of L1 left throttles the loads in flight (measured: 4x slower).  After the last k-slab the
// same 8 warps read the accumulator (tcgen05.ld 32x32b) and apply the epilogue (split-k partial store,
// or beta*C + bias -> relu -> dropout / relu-grad mask).
//
// Shared-memory operand layout: BOTH operands are staged K-major, whatever their layout in global
// memory (MN-major tf32 operands need the special 128B_BASE32B swizzle on this hardware; transposing in
// the loader keeps one well-understood layout).  Slab tile [R rows][16 k] (64 B of K per row), made of
// 8-row x 16-byte core matrices:
//     byte(r, k) = (k/4)*LBO + (r/8)*SBO + (r%8)*16 + (k%4)*4,   SBO = 144, LBO = 18*R + 32
//   SBO = 128 + 16 and LBO = 2 (mod 8) x 16 bytes skew successive row-groups / k-chunks by one / two 16-byte bank
//   groups, so both store patterns below are bank-conflict free:
//     K-contiguous source : a quarter-warp writes the k-chunks of two rows           (16-byte stores)
//     MN-contiguous source: a thread loads a 4(k) x 4(mn) block with four 16-byte loads, transposes it
//                           in registers and writes four 16-byte chunks (rows r..r+3 of one k-chunk);
//                           a quarter-warp covers 8 consecutive 4-row groups.
#include "common.cuh"
#include "tcgen05.cuh"

namespace plnlp {

namespace {

constexpr int TBM = 128;           // CTA tile rows  (UMMA M)
constexpr int TBK = 16;            // k-slab (fp32 elements)
constexpr int KQ = TBK / 4;        // 16-byte k-chunks per row per slab
constexpr int LOADERS = 256;       // threads of warps 0-7
constexpr int NTHREADS = 288;

struct TcGemmParams {
    int64_t M, N, K;
    const float* A; int64_t lda;
    const float* B; int64_t ldb;
    float* C; int64_t ldc;
    float beta;
    const float* bias;
    int act;
    const float* aux; int64_t ldaux;
    float drop_p; uint64_t seed;
    float* ws;
    int split_k;
    int64_t k_per_split;
    int passes;                    // 1 or 3
};

__host__ __device__ constexpr int tile_lbo(int rows) { return 18 * rows + 32; }
constexpr int TILE_SBO = 144;
__host__ __device__ constexpr int slot_bytes(int rows) { return KQ * tile_lbo(rows); }
// 16-byte register chunks a loader thread holds for one operand slab
__host__ __device__ constexpr int nreg(int rows, bool mn) {
    return mn ? 4 * (((rows / 4) * KQ + LOADERS - 1) / LOADERS) : (rows * KQ) / LOADERS;
}

// one 16-byte chunk of an operand tile -> shared memory (hi and, when SPLIT, lo parts)
template <bool SPLIT>
__device__ __forceinline__ void put_chunk(uint8_t* hi, uint8_t* lo, int off, const float (&v)[4]) {
    float h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        h[e] = tc::to_tf32(v[e]);
        l[e] = v[e] - h[e];
    }
    *reinterpret_cast<float4*>(hi + off) = make_float4(h[0], h[1], h[2], h[3]);
    if (SPLIT) *reinterpret_cast<float4*>(lo + off) = make_float4(l[0], l[1], l[2], l[3]);
}

// operand tile loads: plain read-only path.  (L1::no_allocate was measured 6 % SLOWER on B200 for this
// kernel -- 0.94 vs 0.88 ms at 262144x512x512 -- so it is kept only behind a macro.)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
#ifndef PLNLP_GEMM_LDG_NO_ALLOCATE
    r = __ldg(reinterpret_cast<const float4*>(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
#endif
    return r;
}

// Per-thread view of one operand's slabs.  Everything that does not change from slab to slab (row
// pointers, row validity, shared-memory offsets) is computed once; per slab the loader only bumps the
// pointers -- the loader warps are instruction-issue bound, every hoisted instruction counts.
//  MN = false (K-contiguous source, element (r, k) at src[r*ld + k]):
//       chunk c = tid + 256*i: kq = c % KQ, r = c / KQ; reg[i] = 4 consecutive k of row r
//  MN = true  (MN-contiguous source, element (r, k) at src[k*ld + r]):
//       block b = tid + 256*i: rg = b % (R/4), kq = b / (R/4); reg[4*i + j] = rows rg*4..+3 at k = kq*4 + j
template <int R, bool MN, bool VEC>
struct Loader {
    static constexpr int NR = nreg(R, MN);
    static constexpr int NP = MN ? NR / 4 : NR;
    const float* ptr[NP];
    int soff[NP];       // byte offset of the chunk (MN: of row rg*4, rows +1..+3 follow at +16 B)
    int kq4[NP];        // first k of the chunk inside the slab
    int nrow[NP];       // MN scalar path: valid rows of the block (<= 4); otherwise 0/1 row validity
    int64_t step;       // elements between consecutive slabs

    __device__ __forceinline__ void init(const float* src, int64_t ld, int64_t r0, int64_t rows, int64_t kbeg,
                                         int tid) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const int c = tid + LOADERS * i;
            if (!MN) {
                const int kq = c % KQ, r = c / KQ;
                kq4[i] = kq * 4;
                nrow[i] = (r0 + r) < rows ? 1 : 0;
                soff[i] = kq * tile_lbo(R) + (r >> 3) * TILE_SBO + (r & 7) * 16;
                ptr[i] = src + (r0 + r) * ld + kbeg + kq * 4;
            } else {
                const int rg = c % (R / 4), kq = c / (R / 4);
                const bool live = c < (R / 4) * KQ;
                const int64_t left = rows - (r0 + rg * 4);
                kq4[i] = kq * 4;
                nrow[i] = !live ? 0 : (left >= 4 ? 4 : (left > 0 ? static_cast<int>(left) : 0));
                soff[i] = kq * tile_lbo(R) + ((rg * 4) >> 3) * TILE_SBO + ((rg * 4) & 7) * 16;
                ptr[i] = src + (kbeg + kq * 4) * ld + r0 + rg * 4;
            }
        }
        step = MN ? ld * TBK : TBK;
    }

    // kleft = kend - k0 of the slab being fetched (<= 0: nothing left, zero fill)
    __device__ __forceinline__ void fetch(int kleft, float (&reg)[NR][4]) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (!MN) {
                const int lim = kleft - kq4[i];
                if (VEC) {
                    if (nrow[i] && lim > 0) {
                        const float4 t = ld_stream4(ptr[i]);
                        reg[i][0] = t.x; reg[i][1] = t.y; reg[i][2] = t.z; reg[i][3] = t.w;
                    } else {
                        reg[i][0] = reg[i][1] = reg[i][2] = reg[i][3] = 0.0f;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) reg[i][e] = (nrow[i] && e < lim) ? __ldg(ptr[i] + e) : 0.0f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float* q = ptr[i] + j * (step / TBK);
                    float (&d)[4] = reg[4 * i + j];
                    const bool kok = (kq4[i] + j) < kleft;
                    if (VEC) {
                        if (kok && nrow[i]) {
                            const float4 t = ld_stream4(q);
                            d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
                        } else {
                            d[0] = d[1] = d[2] = d[3] = 0.0f;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) d[e] = (kok && e < nrow[i]) ? __ldg(q + e) : 0.0f;
                    }
                }
            }
            ptr[i] += step;
        }
    }

    template <bool SPLIT>
    __device__ __forceinline__ void stash(uint8_t* hi, uint8_t* lo, const float (&reg)[NR][4]) const {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (!MN) {
                put_chunk<SPLIT>(hi, lo, soff[i], reg[i]);
            } else {
                if ((((R / 4) * KQ) % LOADERS != 0) && (threadIdx.x + LOADERS * i >= (R / 4) * KQ)) continue;
#pragma unroll
                for (int e = 0; e < 4; ++e) {      // row rg*4 + e gets (k%4 = 0..3) from the 4 loads
                    const float v[4] = {reg[4 * i + 0][e], reg[4 * i + 1][e], reg[4 * i + 2][e], reg[4 * i + 3][e]};
                    put_chunk<SPLIT>(hi, lo, soff[i] + e * 16, v);
                }
            }
        }
    }
};

__device__ __forceinline__ float tc_epilogue_one(const TcGemmParams& p, int64_t r, int64_t c, float v) {
    if (p.beta != 0.0f) v += p.beta * p.C[r * p.ldc + c];
    if (p.bias) v += __ldg(p.bias + c);
    if (p.act == PLNLP_ACT_RELU) {
        v = fmaxf(v, 0.0f);
        if (p.drop_p > 0.0f)
            v = dropout_keep(p.seed, static_cast<uint64_t>(r) * p.N + c, p.drop_p) ? v * (1.0f / (1.0f - p.drop_p)) : 0.0f;
    } else if (p.act == PLNLP_ACT_RELU_GRAD) {
        v = (__ldg(p.aux + r * p.ldaux + c) > 0.0f) ? v * (1.0f / (1.0f - p.drop_p)) : 0.0f;
    }
    return v;
}

template <bool SPLIT>
struct Cfg {
    static constexpr int STAGES = SPLIT ? 2 : 4;      // 2 CTAs per SM share the 227 KB
};

// BN: CTA tile columns (UMMA N, TMEM columns); AMN/BMN: operand is MN-contiguous in global memory
// (A: transa = 1, B: transb = 0); VA/VB: 16-byte global loads legal; SPLIT: 3xTF32.
template <int BN, bool AMN, bool BMN, bool VA, bool VB, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tcgen05_kernel(const TcGemmParams p) {
    constexpr int STAGES = Cfg<SPLIT>::STAGES;
    constexpr int A_SLOT = slot_bytes(TBM), B_SLOT = slot_bytes(BN);
    constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_SLOT + B_SLOT);
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
    __shared__ uint32_t tmem_holder;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = static_cast<int64_t>(blockIdx.y) * TBM, n0 = static_cast<int64_t>(blockIdx.x) * BN;
    const int64_t kbeg = static_cast<int64_t>(blockIdx.z) * p.k_per_split;
    const int64_t kend = min(p.K, kbeg + p.k_per_split);
    const int n_iter = static_cast<int>((kend - kbeg + TBK - 1) / TBK);
    // columns actually multiplied: N remainder rounded up to the UMMA granularity (16)
    const int64_t n_rem = ((p.N - n0 + 15) / 16) * 16;
    const int n_mma = n_rem < BN ? static_cast<int>(n_rem) : BN;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full_bar[s], LOADERS);
            tc::mbar_init(&empty_bar[s], 1);
        }
        tc::mbar_init(&accum_bar, 1);
        tc::mbar_fence_init();
    }
    if (warp == 8) tc::tmem_alloc<BN>(&tmem_holder);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_holder;

    auto stage_ptr = [&](int s, int which) -> uint8_t* {  // which: 0 A.hi, 1 B.hi, 2 A.lo, 3 B.lo
        uint8_t* base = smem + s * STAGE_BYTES;
        return base + (which & 1 ? A_SLOT : 0) + (which & 2 ? (A_SLOT + B_SLOT) : 0);
    };

    if (warp < 8) {
        // ============================ loaders ============================
        float ra[2][nreg(TBM, AMN)][4], rb[2][nreg(BN, BMN)][4];
        Loader<TBM, AMN, VA> la;
        Loader<BN, BMN, VB> lb;
        la.init(p.A, p.lda, m0, p.M, kbeg, tid);
        lb.init(p.B, p.ldb, n0, p.N, kbeg, tid);
        const int ktot = static_cast<int>(kend - kbeg);
        auto fetch = [&](int it, float (&a)[nreg(TBM, AMN)][4], float (&b)[nreg(BN, BMN)][4]) {
            if (it < n_iter) {                            // called with it = 0, 1, 2, ... in order
                la.fetch(ktot - it * TBK, a);
                lb.fetch(ktot - it * TBK, b);
            }
        };
        auto publish = [&](int it, const float (&a)[nreg(TBM, AMN)][4], const float (&b)[nreg(BN, BMN)][4]) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            tc::mbar_wait(&empty_bar[s], ph ^ 1);         // slot free (first round passes immediately)
            la.template stash<SPLIT>(stage_ptr(s, 0), stage_ptr(s, 2), a);
            lb.template stash<SPLIT>(stage_ptr(s, 1), stage_ptr(s, 3), b);
        };
        fetch(0, ra[0], rb[0]);
        fetch(1, ra[1], rb[1]);
        // NOTE the order: the proxy fence waits for every memory operation the thread has in flight,
        // including global loads, so it must come BEFORE the next prefetch is issued (measured: with
        // the prefetch first every slab paid a full global-load latency).
        for (int it = 0; it < n_iter; it += 2) {
            publish(it, ra[0], rb[0]);
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&full_bar[it % STAGES]);
            fetch(it + 2, ra[0], rb[0]);                  // two slabs ahead of the stores
            if (it + 1 < n_iter) {
                publish(it + 1, ra[1], rb[1]);
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(&full_bar[(it + 1) % STAGES]);
                fetch(it + 3, ra[1], rb[1]);
            }
        }
    } else {
        // ============================ MMA issuer ============================
        const uint32_t idesc = tc::make_idesc_tf32(TBM, n_mma, 0, 0);      // both operands K-major in smem
        constexpr uint32_t A_LBO = tile_lbo(TBM), B_LBO = tile_lbo(BN), A_SBO = TILE_SBO, B_SBO = TILE_SBO;
        // one UMMA consumes K = 8 (32 bytes) = two 16-byte k-chunks, i.e. 2*LBO per step
        constexpr uint32_t A_STEP = 2 * A_LBO, B_STEP = 2 * B_LBO;
        for (int it = 0; it < n_iter; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            tc::mbar_wait(&full_bar[s], ph);
            tc::fence_after_sync();
            if (lane == 0) {
                const uint32_t a_hi = tc::smem_u32(stage_ptr(s, 0)), b_hi = tc::smem_u32(stage_ptr(s, 1));
                const uint32_t a_lo = tc::smem_u32(stage_ptr(s, 2)), b_lo = tc::smem_u32(stage_ptr(s, 3));
#pragma unroll
                for (int j = 0; j < TBK / 8; ++j) {
                    const uint64_t dah = tc::make_smem_desc(a_hi + j * A_STEP, A_LBO, A_SBO);
                    const uint64_t dbh = tc::make_smem_desc(b_hi + j * B_STEP, B_LBO, B_SBO);
                    tc::mma_tf32_ss(tmem_d, dah, dbh, idesc, (it | j) != 0);
                    if (SPLIT) {
                        const uint64_t dal = tc::make_smem_desc(a_lo + j * A_STEP, A_LBO, A_SBO);
                        const uint64_t dbl = tc::make_smem_desc(b_lo + j * B_STEP, B_LBO, B_SBO);
                        tc::mma_tf32_ss(tmem_d, dah, dbl, idesc, 1u);
                        tc::mma_tf32_ss(tmem_d, dal, dbh, idesc, 1u);
                    }
                }
                tc::mma_commit(&empty_bar[s]);                       // frees the slot when these finish
                if (it == n_iter - 1) tc::mma_commit(&accum_bar);    // accumulator complete
            }
            __syncwarp();
        }
    }

    // ============================ epilogue (warps 0-7) ============================
    if (warp < 8) {
        if (n_iter > 0) {
            tc::mbar_wait(&accum_bar, 0);
            tc::fence_after_sync();
        }
        const int q = warp & 3, half = warp >> 2;
        const int64_t r = m0 + q * 32 + lane;
        const bool split = p.split_k > 1;
        const bool plain = p.beta == 0.0f && p.bias == nullptr && p.act == PLNLP_ACT_NONE;
        const float keep_scale = 1.0f / (1.0f - p.drop_p);
        const bool vec_epi = (p.N % 4 == 0) && (p.ldc % 4 == 0) && (reinterpret_cast<uintptr_t>(p.C) % 16 == 0) &&
                             (!p.bias || reinterpret_cast<uintptr_t>(p.bias) % 16 == 0) &&
                             (!p.aux || ((p.ldaux % 4 == 0) && reinterpret_cast<uintptr_t>(p.aux) % 16 == 0));
        float* wsz = split ? p.ws + static_cast<int64_t>(blockIdx.z) * p.M * p.N : nullptr;
        for (int cb = half * (BN / 2); cb < (half + 1) * (BN / 2); cb += 32) {
            if (cb >= n_mma) break;                                   // warp-uniform
            float v[32];
            if (n_iter > 0) {
                tc::tmem_ld_32x32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(cb), v);
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = 0.0f;
            }
            if (r < p.M) {
                const int64_t c0 = n0 + cb;
                if (split) {
                    float* dst = wsz + r * p.N + c0;
                    if ((p.N % 4 == 0) && c0 + 31 < p.N) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            *reinterpret_cast<float4*>(dst + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (c0 + e < p.N) dst[e] = v[e];
                    }
                } else {
                    if (!plain) {
                        if (vec_epi && c0 + 31 < p.N) {
                            // 4 columns at a time: 16-byte loads of C / bias / aux, one Philox block per
                            // group (element r*N + c uses word c%4 of block (r*N + c)/4 -- the same
                            // stream as dropout_keep, a quarter of the hashing)
                            const float* crow = p.C + r * p.ldc + c0;
                            const float* arow = p.aux ? p.aux + r * p.ldaux + c0 : nullptr;
                            const uint64_t ebase = static_cast<uint64_t>(r) * p.N + c0;
#pragma unroll
                            for (int e = 0; e < 32; e += 4) {
                                float x[4] = {v[e], v[e + 1], v[e + 2], v[e + 3]};
                                if (p.beta != 0.0f) {
                                    const float4 c4 = *reinterpret_cast<const float4*>(crow + e);
                                    x[0] += p.beta * c4.x; x[1] += p.beta * c4.y;
                                    x[2] += p.beta * c4.z; x[3] += p.beta * c4.w;
                                }
                                if (p.bias) {
                                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + e));
                                    x[0] += b4.x; x[1] += b4.y; x[2] += b4.z; x[3] += b4.w;
                                }
                                if (p.act == PLNLP_ACT_RELU) {
                                    bool keep[4] = {true, true, true, true};
                                    if (p.drop_p > 0.0f) dropout_keep4(p.seed, ebase + e, p.drop_p, keep);
#pragma unroll
                                    for (int t = 0; t < 4; ++t) x[t] = keep[t] ? fmaxf(x[t], 0.0f) * keep_scale : 0.0f;
                                } else if (p.act == PLNLP_ACT_RELU_GRAD) {
                                    const float4 a4 = __ldg(reinterpret_cast<const float4*>(arow + e));
                                    x[0] = a4.x > 0.0f ? x[0] * keep_scale : 0.0f;
                                    x[1] = a4.y > 0.0f ? x[1] * keep_scale : 0.0f;
                                    x[2] = a4.z > 0.0f ? x[2] * keep_scale : 0.0f;
                                    x[3] = a4.w > 0.0f ? x[3] * keep_scale : 0.0f;
                                }
                                v[e] = x[0]; v[e + 1] = x[1]; v[e + 2] = x[2]; v[e + 3] = x[3];
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if (c0 + e < p.N) v[e] = tc_epilogue_one(p, r, c0 + e, v[e]);
                        }
                    }
                    float* dst = p.C + r * p.ldc + c0;
                    if ((p.ldc % 4 == 0) && (reinterpret_cast<uintptr_t>(p.C) % 16 == 0) && c0 + 31 < p.N) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            *reinterpret_cast<float4*>(dst + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (c0 + e < p.N) dst[e] = v[e];
                    }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc<BN>(tmem_d);
}

__global__ void __launch_bounds__(256) tc_splitk_reduce_kernel(const TcGemmParams p) {
    const int64_t total = p.M * p.N;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float v = 0.0f;
        for (int z = 0; z < p.split_k; ++z) v += p.ws[static_cast<int64_t>(z) * total + i];
        const int64_t r = i / p.N, c = i % p.N;
        p.C[r * p.ldc + c] = tc_epilogue_one(p, r, c, v);
    }
}

template <int BN, bool AMN, bool BMN, bool VA, bool VB, bool SPLIT>
int launch_one(const TcGemmParams& p, dim3 grid, cudaStream_t st) {
    constexpr int bytes = Cfg<SPLIT>::STAGES * (SPLIT ? 2 : 1) * (slot_bytes(TBM) + slot_bytes(BN));
    static_assert(bytes <= 113 * 1024, "two CTAs per SM: epilogue of one overlaps the main loop of the other");
    auto kern = gemm_tcgen05_kernel<BN, AMN, BMN, VA, VB, SPLIT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured = true;
    }
    kern<<<grid, NTHREADS, bytes, st>>>(p);
    return 0;
}

template <int BN, bool AMN, bool BMN>
int launch_vec(const TcGemmParams& p, bool va, bool vb, dim3 grid, cudaStream_t st) {
    // the unaligned (scalar-load) path is only instantiated for the exact 3-pass variant to bound
    // compile time; plain-TF32 requests with unaligned operands use it too
    const bool split = p.passes == 3;
    if (va && vb) {
        return split ? launch_one<BN, AMN, BMN, true, true, true>(p, grid, st)
                     : launch_one<BN, AMN, BMN, true, true, false>(p, grid, st);
    }
    if (va) return launch_one<BN, AMN, BMN, true, false, true>(p, grid, st);
    if (vb) return launch_one<BN, AMN, BMN, false, true, true>(p, grid, st);
    return launch_one<BN, AMN, BMN, false, false, true>(p, grid, st);
}

template <int BN>
int launch_major(const TcGemmParams& p, bool amn, bool bmn, bool va, bool vb, dim3 grid, cudaStream_t st) {
    if (!amn && !bmn) return launch_vec<BN, false, false>(p, va, vb, grid, st);
    if (!amn && bmn) return launch_vec<BN, false, true>(p, va, vb, grid, st);
    if (amn && !bmn) return launch_vec<BN, true, false>(p, va, vb, grid, st);
    return launch_vec<BN, true, true>(p, va, vb, grid, st);
}

}  // namespace
}  // namespace plnlp

extern "C" int plnlp_gemm_tf32(int passes, int transa, int transb, int64_t M, int64_t N, int64_t K, const float* A,
                               int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, float beta,
                               const float* bias, int act, const float* aux, int64_t ldaux, float drop_p,
                               uint64_t seed, float* workspace, int64_t workspace_bytes, int split_k,
                               void* stream) {
    using namespace plnlp;
    PLNLP_REQUIRE(passes == 1 || passes == 3, PLNLP_E_UNSUPPORTED);
    PLNLP_REQUIRE(M >= 0 && N >= 0 && K >= 0, PLNLP_E_SIZE);
    if (M == 0 || N == 0) return 0;
    PLNLP_REQUIRE(A && B && C, PLNLP_E_NULL);
    PLNLP_REQUIRE(lda >= (transa ? M : K) && ldb >= (transb ? K : N) && ldc >= N, PLNLP_E_SIZE);
    PLNLP_REQUIRE(act >= 0 && act <= 2 && drop_p >= 0.0f && drop_p < 1.0f, PLNLP_E_SIZE);
    if (act == PLNLP_ACT_RELU_GRAD) PLNLP_REQUIRE(aux && ldaux >= N, PLNLP_E_NULL);
    if (split_k < 1 || K == 0) split_k = 1;
    TcGemmParams p{};
    p.M = M; p.N = N; p.K = K; p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
    p.beta = beta; p.bias = bias; p.act = act; p.aux = aux; p.ldaux = ldaux; p.drop_p = drop_p; p.seed = seed;
    p.passes = passes;
    int64_t kper = ceil_div(ceil_div(K, split_k), TBK) * TBK;
    if (kper == 0) kper = TBK;
    p.k_per_split = kper;
    p.split_k = split_k = static_cast<int>(K == 0 ? 1 : ceil_div(K, kper));
    p.ws = workspace;
    if (split_k > 1) {
        PLNLP_REQUIRE(workspace, PLNLP_E_NULL);
        PLNLP_REQUIRE(workspace_bytes >= static_cast<int64_t>(split_k) * M * N * 4, PLNLP_E_WORKSPACE);
        PLNLP_REQUIRE(aligned(workspace, 16), PLNLP_E_ALIGN);
    }
    const bool amn = transa != 0, bmn = transb == 0;
    const bool va = aligned(A, 16) && (lda % 4 == 0) && ((transa ? M : K) % 4 == 0);
    const bool vb = aligned(B, 16) && (ldb % 4 == 0) && ((transb ? K : N) % 4 == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if (N > 128) {
        const dim3 grid(static_cast<unsigned>(ceil_div(N, 256)), static_cast<unsigned>(ceil_div(M, TBM)),
                        static_cast<unsigned>(split_k));
        rc = launch_major<256>(p, amn, bmn, va, vb, grid, st);
    } else {
        const dim3 grid(static_cast<unsigned>(ceil_div(N, 128)), static_cast<unsigned>(ceil_div(M, TBM)),
                        static_cast<unsigned>(split_k));
        rc = launch_major<128>(p, amn, bmn, va, vb, grid, st);
    }
    if (rc != 0) return rc;
    PLNLP_LAUNCH_CHECK();
    if (split_k > 1) {
        const int64_t total = M * N;
        const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div(total, 256), 148 * 8));
        tc_splitk_reduce_kernel<<<blocks, 256, 0, st>>>(p);
        PLNLP_LAUNCH_CHECK();
    }
    return 0;
}
