// Shared pieces of the tcgen05 TF32 GEMM kernels (gemm_tcgen05.cu: one CTA per tile, cta_group::1;
// gemm_tcgen05_2cta.cu: CTA pairs, cta_group::2): parameters, the K-major shared-memory tile layout, the
// loader that splits fp32 operands into tf32 hi/lo parts, and the scalar epilogue.  See gemm_tcgen05.cu
// for the layout description.
#pragma once
#include "common.cuh"
#include "tcgen05.cuh"

namespace plnlp {
namespace tcgemm {


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
    // fused edge scoring (csrc/gemm_tcgen05_2cta.cu, plnlp_edge_mlp_fwd_tf32): when a_edges != NULL the A
    // operand row p is the Hadamard product h[src_p, :] * h[dst_p, :] gathered on the fly (A = h, lda = ldh,
    // a_rows = rows of h); when w_out != NULL the epilogue also accumulates score_part[q][r] = sum over
    // this CTA's column half of C[r, c] * w_out[c]; C == NULL skips the store of C.
    const int64_t* a_edges; int64_t a_rows;
    const float* w_out; float* score_part; int64_t score_ld;
    // fused edge scoring, backward half (plnlp_edge_mlp_bwd_tf32): when dz_dscore != NULL the A operand is
    //     dZ1[p, c] = A[p, c] > 0 ? (dscore[p] * dz_scale) * w2[c] : 0          (A = the stored activation a1)
    // formed by the loader (DzLoader) instead of being read from HBM -- K-major for dA0 = dZ1 @ W1 (row = pair,
    // k = hidden column), MN-major for dW1 = dZ1^T @ a0 (row = hidden column, k = pair).  When b_edges != NULL the
    // B operand of the second product, a0[p, j] = h[src_p, j] * h[dst_p, j] (B = h, ldb = ldh, b_rows = rows of
    // h), is re-gathered by the loader (HadamardLoaderMN): neither dZ1 nor the Hadamard product exists in HBM.
    const float* dz_dscore; const float* dz_w2; float dz_scale;
    const int64_t* b_edges; int64_t b_rows;
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

// A-operand provider of the fused edge-scoring kernel: slab tile row r is the Hadamard product of the two
// endpoint embeddings of pair (r0 + r), gathered from h (K-contiguous, same chunk mapping and shared-memory
// layout as Loader<R, false, VEC>).  Replaces h[edge[0]] * h[edge[1]] (model.py:155-156, layer.py:81)
// without the [P, H] product ever existing in HBM.
template <int R, bool VEC>
struct GatherLoader {
    static constexpr int NR = nreg(R, false);
    const float* ps[NR];
    const float* pd[NR];
    int soff[NR];
    int kq4[NR];
    int nrow[NR];

    __device__ __forceinline__ void init(const float* h, int64_t ldh, int64_t h_rows, const int64_t* edges,
                                         int64_t r0, int64_t rows, int64_t kbeg, int tid) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int c = tid + LOADERS * i;
            const int kq = c % KQ, r = c / KQ;
            const int64_t pr = r0 + r;
            kq4[i] = kq * 4;
            nrow[i] = pr < rows ? 1 : 0;
            int64_t s = 0, d = 0;
            if (nrow[i]) {
                s = __ldg(edges + 2 * pr);
                d = __ldg(edges + 2 * pr + 1);
                if (s < 0) s += h_rows;
                if (d < 0) d += h_rows;
            }
            soff[i] = kq * tile_lbo(R) + (r >> 3) * TILE_SBO + (r & 7) * 16;
            ps[i] = h + s * ldh + kbeg + kq * 4;
            pd[i] = h + d * ldh + kbeg + kq * 4;
        }
    }

    __device__ __forceinline__ void fetch(int kleft, float (&reg)[NR][4]) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int lim = kleft - kq4[i];
            if (VEC) {
                if (nrow[i] && lim > 0) {
                    const float4 a = ld_stream4(ps[i]), b = ld_stream4(pd[i]);
                    reg[i][0] = a.x * b.x; reg[i][1] = a.y * b.y; reg[i][2] = a.z * b.z; reg[i][3] = a.w * b.w;
                } else {
                    reg[i][0] = reg[i][1] = reg[i][2] = reg[i][3] = 0.0f;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    reg[i][e] = (nrow[i] && e < lim) ? __ldg(ps[i] + e) * __ldg(pd[i] + e) : 0.0f;
            }
            ps[i] += TBK;
            pd[i] += TBK;
        }
    }

    template <bool SPLIT>
    __device__ __forceinline__ void stash(uint8_t* hi, uint8_t* lo, const float (&reg)[NR][4]) const {
#pragma unroll
        for (int i = 0; i < NR; ++i) put_chunk<SPLIT>(hi, lo, soff[i], reg[i]);
    }
};

// A-operand provider of the fused edge-scoring BACKWARD: the stored hidden activation a1 = dropout(relu(z1)) is read
// with Loader's own chunk mapping and turned into dZ1 = (dscore (x) w2) . mask(a1) on the fly -- replaces the
// [P, N1] matrix plnlp_mlp_out_bwd_f32 writes and both GEMMs read back (layer.py:80-87 backward).
//   MN = false (dA0 = dZ1 @ W1):   chunk = 4 consecutive hidden columns of one pair: dscore fixed per chunk slot,
//                                  w2 moves with the slab
//   MN = true  (dW1 = dZ1^T @ a0): block = 4 hidden columns x 4 pairs: w2 fixed per block slot, dscore moves with
//                                  the slab (one 16-byte load)
template <int R, bool MN, bool VEC>
struct DzLoader {
    using Base = Loader<R, MN, VEC>;
    static constexpr int NR = Base::NR;
    static constexpr int NP = Base::NP;
    Base base;
    float fix[NP][4];          // MN: w2 of the block's 4 rows; !MN: fix[i][0] = scaled dscore of the chunk's row
    const float* mov[NP];      // MN: dscore at the block's first k; !MN: w2 at the chunk's first k
    float scale;

    __device__ __forceinline__ void init(const float* a1, int64_t ld, int64_t r0, int64_t rows, int64_t kbeg,
                                         const float* dscore, const float* w2, float s, int tid) {
        base.init(a1, ld, r0, rows, kbeg, tid);
        scale = s;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const int c = tid + LOADERS * i;
            if (!MN) {
                const int r = c / KQ;
                fix[i][0] = (r0 + r) < rows ? __ldg(dscore + r0 + r) * s : 0.0f;
                fix[i][1] = fix[i][2] = fix[i][3] = 0.0f;
                mov[i] = w2 + kbeg + base.kq4[i];
            } else {
                const int rg = c % (R / 4);
#pragma unroll
                for (int e = 0; e < 4; ++e) fix[i][e] = (r0 + rg * 4 + e) < rows ? __ldg(w2 + r0 + rg * 4 + e) : 0.0f;
                mov[i] = dscore + kbeg + base.kq4[i];
            }
        }
    }

    __device__ __forceinline__ void fetch(int kleft, float (&reg)[NR][4]) {
        base.fetch(kleft, reg);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            float m[4];
            if (base.kq4[i] + 3 < kleft) {                 // whole group inside K: one 16-byte load
                const float4 t = __ldg(reinterpret_cast<const float4*>(mov[i]));
                m[0] = t.x; m[1] = t.y; m[2] = t.z; m[3] = t.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) m[e] = (base.kq4[i] + e) < kleft ? __ldg(mov[i] + e) : 0.0f;
            }
            if (!MN) {
#pragma unroll
                for (int e = 0; e < 4; ++e) reg[i][e] = reg[i][e] > 0.0f ? fix[i][0] * m[e] : 0.0f;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float d = m[j] * scale;
#pragma unroll
                    for (int e = 0; e < 4; ++e) reg[4 * i + j][e] = reg[4 * i + j][e] > 0.0f ? d * fix[i][e] : 0.0f;
                }
            }
            mov[i] += TBK;
        }
    }

    template <bool SPLIT>
    __device__ __forceinline__ void stash(uint8_t* hi, uint8_t* lo, const float (&reg)[NR][4]) const {
        base.template stash<SPLIT>(hi, lo, reg);
    }
};

// B-operand provider of dW1 = dZ1^T @ a0: element (n = column j of h, k = pair p) is h[src_p, j] * h[dst_p, j], gathered
// with Loader<R, true, VEC>'s block mapping (4 columns x 4 pairs per block).  The endpoint ids of the NEXT slab are
// fetched while the current slab's rows are in flight, so a fetch never waits on an index load.
template <int R, bool VEC>
struct HadamardLoaderMN {
    static constexpr int NR = nreg(R, true);
    static constexpr int NP = NR / 4;
    static_assert(NP == 1, "one block per loader thread");
    const float* hcol;          // h + first column of the block
    const int64_t* eptr;        // edges of the NEXT slab's first pair of this block
    int64_t ldh, h_rows;
    int es[4], ed[4];           // endpoint rows of the current slab's 4 pairs (h has < 2^31 rows)
    int soff, kq4, nrow;
    bool live;

    __device__ __forceinline__ void load_ids(int kleft) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t s = 0, d = 0;
            if (live && (kq4 + j) < kleft) {
                const longlong2 sd = __ldg(reinterpret_cast<const longlong2*>(eptr) + j);     // (src, dst) of one pair
                s = sd.x < 0 ? sd.x + h_rows : sd.x;
                d = sd.y < 0 ? sd.y + h_rows : sd.y;
            }
            es[j] = static_cast<int>(s);
            ed[j] = static_cast<int>(d);
        }
        eptr += 2 * TBK;
    }

    __device__ __forceinline__ void init(const float* h, int64_t ld, int64_t rows_h, const int64_t* edges, int64_t n0,
                                         int64_t n_end, int64_t kbeg, int ktot, int tid) {
        const int rg = tid % (R / 4), kq = tid / (R / 4);
        live = tid < (R / 4) * KQ;
        const int64_t left = n_end - (n0 + rg * 4);
        kq4 = kq * 4;
        nrow = !live ? 0 : (left >= 4 ? 4 : (left > 0 ? static_cast<int>(left) : 0));
        soff = kq * tile_lbo(R) + ((rg * 4) >> 3) * TILE_SBO + ((rg * 4) & 7) * 16;
        hcol = h + n0 + rg * 4;
        ldh = ld;
        h_rows = rows_h;
        eptr = edges + 2 * (kbeg + kq4);
        load_ids(ktot);
    }

    __device__ __forceinline__ void fetch(int kleft, float (&reg)[NR][4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float (&d)[4] = reg[j];
            const bool kok = (kq4 + j) < kleft;
            const float* ps = hcol + static_cast<int64_t>(es[j]) * ldh;
            const float* pd = hcol + static_cast<int64_t>(ed[j]) * ldh;
            if (VEC) {
                if (kok && nrow) {
                    const float4 a = ld_stream4(ps), b = ld_stream4(pd);
                    d[0] = a.x * b.x; d[1] = a.y * b.y; d[2] = a.z * b.z; d[3] = a.w * b.w;
                } else {
                    d[0] = d[1] = d[2] = d[3] = 0.0f;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) d[e] = (kok && e < nrow) ? __ldg(ps + e) * __ldg(pd + e) : 0.0f;
            }
        }
        load_ids(kleft - TBK);          // ids of the slab the NEXT call fetches
    }

    template <bool SPLIT>
    __device__ __forceinline__ void stash(uint8_t* hi, uint8_t* lo, const float (&reg)[NR][4]) const {
        if (!live) return;
#pragma unroll
        for (int e = 0; e < 4; ++e) {      // row rg*4 + e gets (k%4 = 0..3) from the 4 loads
            const float v[4] = {reg[0][e], reg[1][e], reg[2][e], reg[3][e]};
            put_chunk<SPLIT>(hi, lo, soff + e * 16, v);
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


// beta*C + bias -> relu -> dropout / relu-grad mask for the 32 columns c0 .. c0+31 of row r held in v
// (bias_s: optional copy of the bias vector in shared memory, indexed by global column -- gemm_tma.cu)
__device__ __forceinline__ void tc_epi_apply32(const TcGemmParams& p, int64_t r, int64_t c0, float (&v)[32], bool vec_epi,
                                               float keep_scale, const float* bias_s = nullptr) {
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
                const float4 b4 = bias_s ? *reinterpret_cast<const float4*>(bias_s + c0 + e)
                                         : __ldg(reinterpret_cast<const float4*>(p.bias + c0 + e));
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


// Epilogue of one CTA tile, executed by warps 0-7 after the accumulator barrier: thread (warp q = w%4,
// lane) owns accumulator row q*32 + lane; warps 0-3 / 4-7 take the two column halves.  Split-k partial
// store, or beta*C + bias -> relu -> dropout / relu-grad mask.
// HALVES = 1: four warps (one per TMEM lane quarter) each take all BN columns (gemm_tma.cu).
template <int BN, int HALVES = 2>
__device__ __forceinline__ void tc_epilogue_tile(const TcGemmParams& p, uint32_t tmem_d, int64_t m0, int64_t n0,
                                                 int n_mma, int n_iter, int warp, int lane) {
    const int q = warp & 3, half = HALVES == 2 ? (warp >> 2) : 0;
    const int64_t r = m0 + q * 32 + lane;
    const bool split = p.split_k > 1;
    const bool plain = p.beta == 0.0f && p.bias == nullptr && p.act == PLNLP_ACT_NONE;
    const float keep_scale = 1.0f / (1.0f - p.drop_p);
    float score_part = 0.0f;
    const bool vec_epi = (p.N % 4 == 0) && (!p.C || ((p.ldc % 4 == 0) && (reinterpret_cast<uintptr_t>(p.C) % 16 == 0))) &&
                         (!p.bias || reinterpret_cast<uintptr_t>(p.bias) % 16 == 0) &&
                         (!p.aux || ((p.ldaux % 4 == 0) && reinterpret_cast<uintptr_t>(p.aux) % 16 == 0));
    float* wsz = split ? p.ws + static_cast<int64_t>(blockIdx.z) * p.M * p.N : nullptr;
    for (int cb = half * (BN / HALVES); cb < (half + 1) * (BN / HALVES); cb += 32) {
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
                if (!plain) tc_epi_apply32(p, r, c0, v, vec_epi, keep_scale);
                if (p.w_out) {     // fused out_channels = 1 layer: this thread's share of a[r, :] . w_out
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (c0 + e < p.N) score_part = fmaf(v[e], __ldg(p.w_out + c0 + e), score_part);
                }
                if (p.C) {
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
    // partial q = (column tile, column half): the caller adds the 2*ceil(N/BN) partials in index order
    if (p.w_out && !split && r < p.M)
        p.score_part[(static_cast<int64_t>(blockIdx.y) * 2 + half) * p.score_ld + r] = score_part;
}

// split-k partial reduction + epilogue (defined in gemm_tcgen05.cu)
int tc_splitk_reduce(const TcGemmParams& p, cudaStream_t st);

}  // namespace tcgemm
}  // namespace plnlp
