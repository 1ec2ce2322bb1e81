// CSR row-gather SpMM for sm_100a: out[r,:] = epi( (sum_{p in row r} val[p] * x[col[p],:]) / row_div[r] )
//
// Replaces torch_sparse.matmul under SAGEConv / GCNConv (/root/reference/plnlp/layer.py:20,23)
// and its backward (same kernel on the transposed structure).
//
// Design (HBM/L2-bound gather, no tensor cores):
//  * one warp per work item (a whole row, or a <=chunk slice of a hub row -- see the plan in
//    plnlp_b200/graph.py); lanes own fixed feature columns, so each output element is
//    accumulated strictly in CSR order (bit-identical to the in-order CPU loop for unsplit rows).
//  * column indices are fetched 32 at a time (one coalesced 128 B load) and broadcast by shuffle;
//  * feature rows are gathered with 16-byte loads, NB neighbours x U vectors per lane in flight
//    (8 independent 16 B loads per lane) before any of them is consumed;
//  * valued accumulation uses separate fp32 multiply and add (no FMA contraction) so that the
//    result equals the reference CPU loop bit for bit; the kernel is memory-bound, the extra
//    instruction is free.
#include <cstdlib>

#include <cuda_bf16.h>

#include "common.cuh"

namespace plnlp {

// T = feature / output element type: float, or __nv_bfloat16 (bf16 storage, fp32 accumulation)
template <typename T>
struct SpmmParamsT {
    const int32_t* item_ptr;
    const int32_t* item_row;
    const int32_t* item_slot;
    int64_t n_items;
    const int32_t* item_end;   // optional explicit item ends (row-subset plans); NULL: item i ends at item_ptr[i+1]
    const int32_t* x_index;    // optional: source j reads row x_index[j] of x; x_index[j] < 0 = an all-zero row (skipped)
    const int32_t* col;
    const float* val;
    const float* row_div;
    const float* bias;
    int relu;
    float drop_p;
    uint64_t seed;
    const T* x;
    int64_t ldx;
    T* out;
    int64_t ldo;
    int F;
    float* partial;
    const int32_t* fix_ptr;
    const int32_t* fix_row;
    int64_t n_fix;
    // optional relu(-dropout) backward mask fused into the epilogue: out = mask[row, f] > 0 ? out * mask_scale : 0.
    // `mask` is the forward activation Y = dropout(relu(.)) whose gradient this product is (the backward SpMM of the
    // conv that consumed Y then hands the previous layer the gradient w.r.t. its PRE-activation: no separate
    // relu-backward pass over [N, F])
    const float* mask;
    int64_t ldmask;
    float mask_scale;
    // L2 prefetch of the gathered rows (see prefetch_rows): 0 off, 1 one prefetch.global.L2 per 128-byte line,
    // 2 one cp.async.bulk.prefetch.L2 per row
    int pf;
};
using SpmmParams = SpmmParamsT<float>;

// process-wide tuning knobs (plnlp_spmm_tune): L2 prefetch mode of the gather kernels, shared-memory staged kernel
// configuration for narrow fp32 rows (0 = off), warps per CTA of the staged kernel
static int g_spmm_pf = 3;
static int g_spmm_staged = 12;
static int g_spmm_staged_warps = 4;
// mode 3: the bulk prefetch where it was measured to pay (profiles/r02_spmm_tune_ab.txt, citation2-shape graph): fp32
// rows of 257 .. 704 bytes (F = 96: 68 % -> 84 % of the HBM copy peak, 128: 83 -> 90, 160: 74 -> 81; F = 200 and 256
// lose 2 - 10 %, rows of <= 256 bytes run on the staged kernel), bf16 rows of 256 .. 1024 bytes (F = 128: 57 -> 61,
// 256: 65 -> 80, 512: 75 -> 86)
static inline int resolve_pf(int64_t row_bytes, int elem) {
    if (g_spmm_pf != 3) return g_spmm_pf;
    if (elem == 4) return (row_bytes > 256 && row_bytes <= 704) ? 2 : 0;
    return (row_bytes >= 256 && row_bytes <= 1024) ? 2 : 0;
}

// The gather kernels are LATENCY bound: a warp walks its row NB entries at a time and every step waits a full DRAM
// round trip (ncu: > 80 % of the stall samples are long-scoreboard waits on the gathered rows), registers cap the
// rows in flight at ~256 per SM.  As soon as the (up to 32) column indices of a batch are known, lane j asks the L2
// for row col[j]: every row of the batch is in flight at once, costs no register, and the gather loop that follows
// finds its rows in (or on their way to) the L2.
//   mode 1: prefetch.global.L2 for every 128-byte line the slice [first, first + bytes) of the row touches
//   mode 2: ONE cp.async.bulk.prefetch.L2 for the 16-byte aligned INSIDE of the slice (the instruction wants aligned
//           address and size; the window never leaves the row), plus a line prefetch for an unaligned head / tail
template <typename T>
__device__ __forceinline__ void prefetch_row(int mode, const T* first, int bytes) {
    const uintptr_t b0 = reinterpret_cast<uintptr_t>(first), b1 = b0 + bytes;
    if (mode == 1) {
        for (uintptr_t a = b0 & ~static_cast<uintptr_t>(127); a < b1; a += 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    } else {
        const uintptr_t a0 = (b0 + 15) & ~static_cast<uintptr_t>(15), a1 = b1 & ~static_cast<uintptr_t>(15);
        if (a1 > a0) {
            const uint32_t sz = static_cast<uint32_t>(a1 - a0);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(sz) : "memory");
        }
        if (a0 != b0 || a1 <= a0) asm volatile("prefetch.global.L2 [%0];" ::"l"(b0));
        if (a1 != b1 && a1 > a0) asm volatile("prefetch.global.L2 [%0];" ::"l"(b1 - 1));
    }
}

// bf16 rows: VEC elements per lane (16 / 8 / 4 / 2 bytes), widened to fp32 in registers
template <int VEC>
__device__ __forceinline__ void load_vec(float (&d)[VEC], const __nv_bfloat16* p) {
    if constexpr (VEC == 8) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            d[2 * i] = __uint_as_float(w[i] << 16);
            d[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    } else if constexpr (VEC == 4) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        d[0] = __uint_as_float(t.x << 16); d[1] = __uint_as_float(t.x & 0xffff0000u);
        d[2] = __uint_as_float(t.y << 16); d[3] = __uint_as_float(t.y & 0xffff0000u);
    } else if constexpr (VEC == 2) {
        const uint32_t t = __ldg(reinterpret_cast<const uint32_t*>(p));
        d[0] = __uint_as_float(t << 16); d[1] = __uint_as_float(t & 0xffff0000u);
    } else {
        d[0] = __bfloat162float(p[0]);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec(__nv_bfloat16* p, const float (&d)[VEC]) {
    if constexpr (VEC == 1) {
        p[0] = __float2bfloat16_rn(d[0]);
    } else {
        uint32_t w[VEC / 2];
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(d[2 * i], d[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        if constexpr (VEC == 8) *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
        else if constexpr (VEC == 4) *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
        else *reinterpret_cast<uint32_t*>(p) = w[0];
    }
}

// finished-row epilogue for the VEC features starting at column f of row `row`
template <typename T, int VEC>
__device__ __forceinline__ void finish_store(const SpmmParamsT<T>& p, int row, int f, float (&a)[VEC]) {
    if (p.row_div) {
        const float d = __ldg(p.row_div + row);
#pragma unroll
        for (int e = 0; e < VEC; ++e) a[e] = a[e] / d;
    }
    if (p.bias) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) a[e] += __ldg(p.bias + f + e);
    }
    if (p.relu) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) a[e] = fmaxf(a[e], 0.0f);
    }
    if (p.drop_p > 0.0f) {
        const float s = 1.0f / (1.0f - p.drop_p);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const uint64_t idx = static_cast<uint64_t>(row) * static_cast<uint64_t>(p.F) + (f + e);
            a[e] = dropout_keep(p.seed, idx, p.drop_p) ? a[e] * s : 0.0f;
        }
    }
    if (p.mask) {
        const float* m = p.mask + static_cast<int64_t>(row) * p.ldmask + f;
#pragma unroll
        for (int e = 0; e < VEC; ++e) a[e] = __ldg(m + e) > 0.0f ? a[e] * p.mask_scale : 0.0f;
    }
    store_vec<VEC>(p.out + static_cast<int64_t>(row) * p.ldo + f, a);
}

template <typename T, int VEC, int U, int NB, bool HAS_VAL, bool PRED>
__device__ __forceinline__ void gather_block(const T* __restrict__ xb, int64_t ldx, int c, float v,
                                             int j, int n, const bool (&act)[U], float (&acc)[U][VEC]) {
    float t[NB][U][VEC];
    float vv[NB];
    bool ok[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int jj = j + b;
        const int cj = __shfl_sync(0xffffffffu, c, jj & 31);
        vv[b] = HAS_VAL ? __shfl_sync(0xffffffffu, v, jj & 31) : 1.0f;
        ok[b] = !PRED || (jj < n);
        const T* row = xb + static_cast<int64_t>(cj) * ldx;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (ok[b] && act[u]) {
                load_vec<VEC>(t[b][u], row + u * 32 * VEC);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) t[b][u][e] = 0.0f;
            }
        }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (ok[b]) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    acc[u][e] = HAS_VAL ? __fadd_rn(acc[u][e], __fmul_rn(vv[b], t[b][u][e]))
                                        : __fadd_rn(acc[u][e], t[b][u][e]);
        }
    }
}

// Row-sparse operand: only the entries whose source row is flagged non-zero are gathered.  `live` is the warp's
// ballot over the current 32 entries; the next NB set bits (in entry order, so the accumulation order of the
// surviving terms is unchanged and the skipped terms are exact zeros) are consumed per call.
template <typename T, int VEC, int U, int NB, bool HAS_VAL>
__device__ __forceinline__ void gather_block_masked(const T* __restrict__ xb, int64_t ldx, int c, float v,
                                                    unsigned& live, const bool (&act)[U], float (&acc)[U][VEC]) {
    float t[NB][U][VEC];
    float vv[NB];
    bool ok[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        ok[b] = live != 0u;
        const int src = ok[b] ? (__ffs(live) - 1) : 0;
        live &= live - 1u;                                   // 0 stays 0
        const int cj = __shfl_sync(0xffffffffu, c, src);
        vv[b] = HAS_VAL ? __shfl_sync(0xffffffffu, v, src) : 1.0f;
        const T* row = xb + static_cast<int64_t>(cj) * ldx;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (ok[b] && act[u]) {
                load_vec<VEC>(t[b][u], row + u * 32 * VEC);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) t[b][u][e] = 0.0f;
            }
        }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (ok[b]) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    acc[u][e] = HAS_VAL ? __fadd_rn(acc[u][e], __fmul_rn(vv[b], t[b][u][e]))
                                        : __fadd_rn(acc[u][e], t[b][u][e]);
        }
    }
}

// NB = neighbours whose rows are in flight together (NB x U 16-byte loads per lane)
// The kernel is latency bound (ncu: >80 % of the stall samples are long-scoreboard waits on the gathers), so
// resident warps matter more than per-warp depth: variants whose in-flight tile fits 8 registers are held to
// 32 registers per thread (8 blocks = all 64 warp slots of the SM), 16-register tiles to 40 (6 blocks).
template <typename T, int VEC, int U, int NB, bool HAS_VAL>
__global__ void __launch_bounds__(256, (NB * U * VEC <= 8) ? 8 : (NB * U * VEC <= 16) ? 6 : (sizeof(T) == 2) ? 4 : 3)
    spmm_csr_kernel(const SpmmParamsT<T> p) {
    const int lane = threadIdx.x & 31;
    const int64_t item = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= p.n_items) return;
    const int fbase = blockIdx.y * (32 * VEC * U) + lane * VEC;
    bool act[U];
#pragma unroll
    for (int u = 0; u < U; ++u) act[u] = (fbase + u * 32 * VEC) < p.F;

    float acc[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[u][e] = 0.0f;

    const int beg = __ldg(p.item_ptr + item);
    const int end = p.item_end ? __ldg(p.item_end + item) : __ldg(p.item_ptr + item + 1);
    const T* __restrict__ xb = p.x + fbase;
    // this CTA column's slice of a source row (what the prefetch asks for)
    const int slab0 = blockIdx.y * (32 * VEC * U);
    const T* __restrict__ xrow0 = p.x + slab0;
    const int pf_bytes = static_cast<int>(sizeof(T)) * min(32 * VEC * U, p.F - slab0);
    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int c = 0;
        float v = 0.0f;
        if (lane < n) {
            c = __ldg(p.col + base + lane);
            if (HAS_VAL) v = __ldg(p.val + base + lane);
        }
        if (p.x_index) {                                         // warp-uniform
            if (lane < n) c = __ldg(p.x_index + c);
            if (p.pf && lane < n && c >= 0) prefetch_row(p.pf, xrow0 + static_cast<int64_t>(c) * p.ldx, pf_bytes);
            unsigned live = __ballot_sync(0xffffffffu, lane < n && c >= 0);
#pragma unroll 1
            while (live) gather_block_masked<T, VEC, U, NB, HAS_VAL>(xb, p.ldx, c, v, live, act, acc);
        } else if (n == 32) {
            if (p.pf) prefetch_row(p.pf, xrow0 + static_cast<int64_t>(c) * p.ldx, pf_bytes);
#pragma unroll 1
            for (int j = 0; j < 32; j += NB)
                gather_block<T, VEC, U, NB, HAS_VAL, false>(xb, p.ldx, c, v, j, n, act, acc);
        } else {
            if (p.pf && lane < n) prefetch_row(p.pf, xrow0 + static_cast<int64_t>(c) * p.ldx, pf_bytes);
#pragma unroll 1
            for (int j = 0; j < n; j += NB)
                gather_block<T, VEC, U, NB, HAS_VAL, true>(xb, p.ldx, c, v, j, n, act, acc);
        }
    }

    const int row = __ldg(p.item_row + item);
    const int slot = __ldg(p.item_slot + item);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (!act[u]) continue;
        const int f = fbase + u * 32 * VEC;
        if (slot >= 0) {
            store_vec<VEC>(p.partial + static_cast<int64_t>(slot) * p.F + f, acc[u]);
        } else {
            finish_store<T, VEC>(p, row, f, acc[u]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Narrow rows (F <= 64 floats on a 16-byte aligned pitch): TWO work items per warp.
//
// With one warp per row a 50-wide row keeps 25 lanes busy with 8-byte loads and 800 B in flight per warp; the
// kernel is latency bound (see above), so rows this short run at ~50 % of the HBM roof.  Here each half-warp owns
// one item: 16 lanes x one 16-byte load cover a row of up to 64 floats (the last vector of a row whose width is not
// a multiple of 4 reads pitch padding, which is never stored), NB neighbours per half are in flight, so a warp
// keeps 2 x NB x 16 B x 16 lanes in flight with the same register footprint.  Accumulation order per output
// element is still the CSR order of its row (bit-identical results).  Measured on the citation2-shape graph:
// F = 64 5.07 -> 3.78 ms (53 % -> 71 % of the HBM copy peak), F = 32 4.63 -> 3.38 ms.  Below ~256 B per gathered
// row the time no longer falls with the width (3.0 ms at F = 16): the bound is the rate of random DRAM row
// fetches, not bytes.
// ---------------------------------------------------------------------------------------------------------
template <typename T, bool HAS_VAL, int NB>
__global__ void __launch_bounds__(256, (sizeof(T) == 4) ? 8 : 6) spmm_csr_narrow_kernel(const SpmmParamsT<T> p) {
    constexpr int VEC = 16 / sizeof(T);                 // elements per 16-byte load: 4 fp32 / 8 bf16
    const int lane = threadIdx.x & 31;
    const int half = lane >> 4, hl = lane & 15;
    const int64_t warp_id = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t item = warp_id * 2 + half;
    const bool valid = item < p.n_items;
    const int f = hl * VEC;
    const bool act = f < p.F;

    int beg = 0, end = 0;
    if (valid) {
        beg = __ldg(p.item_ptr + item);
        end = p.item_end ? __ldg(p.item_end + item) : __ldg(p.item_ptr + item + 1);
    }
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.0f;
    const T* __restrict__ xb = p.x + f;
    const int src0 = half << 4;                        // first lane of this half
    for (int base = beg; __any_sync(0xffffffffu, base < end); base += 16) {
        const int n = max(0, min(16, end - base));
        int c = 0;
        float v = 0.0f;
        if (hl < n) {
            c = __ldg(p.col + base + hl);
            if (HAS_VAL) v = __ldg(p.val + base + hl);
        }
        if (p.pf && hl < n) prefetch_row(p.pf, p.x + static_cast<int64_t>(c) * p.ldx, p.F * static_cast<int>(sizeof(T)));
        const int n_other = __shfl_xor_sync(0xffffffffu, n, 16);
        const int nmax = max(n, n_other);              // warp-uniform trip count
#pragma unroll 1
        for (int j = 0; j < nmax; j += NB) {
            float t[NB][VEC];
            float vv[NB];
            bool ok[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int jj = j + b;
                const int cj = __shfl_sync(0xffffffffu, c, src0 | (jj & 15));
                vv[b] = HAS_VAL ? __shfl_sync(0xffffffffu, v, src0 | (jj & 15)) : 1.0f;
                ok[b] = jj < n;
                if (ok[b] && act) {
                    load_vec<VEC>(t[b], xb + static_cast<int64_t>(cj) * p.ldx);
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) t[b][e] = 0.0f;
                }
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if (ok[b]) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e)
                        acc[e] = HAS_VAL ? __fadd_rn(acc[e], __fmul_rn(vv[b], t[b][e])) : __fadd_rn(acc[e], t[b][e]);
                }
            }
        }
    }
    if (!valid || !act) return;
    const int row = __ldg(p.item_row + item);
    const int slot = __ldg(p.item_slot + item);
    const int live = min(VEC, p.F - f);                // columns of this lane that exist
    if (slot >= 0) {
        float* dst = p.partial + static_cast<int64_t>(slot) * p.F + f;
        for (int e = 0; e < live; ++e) dst[e] = acc[e];
        return;
    }
    if (p.row_div) {
        const float d = __ldg(p.row_div + row);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = acc[e] / d;
    }
    if (p.bias) {
        for (int e = 0; e < live; ++e) acc[e] += __ldg(p.bias + f + e);
    }
    if (p.relu) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = fmaxf(acc[e], 0.0f);
    }
    if (p.drop_p > 0.0f) {
        const float s = 1.0f / (1.0f - p.drop_p);
        for (int e = 0; e < live; ++e) {
            const uint64_t idx = static_cast<uint64_t>(row) * static_cast<uint64_t>(p.F) + (f + e);
            acc[e] = dropout_keep(p.seed, idx, p.drop_p) ? acc[e] * s : 0.0f;
        }
    }
    if (p.mask) {
        const float* m = p.mask + static_cast<int64_t>(row) * p.ldmask + f;
        for (int e = 0; e < live; ++e) acc[e] = __ldg(m + e) > 0.0f ? acc[e] * p.mask_scale : 0.0f;
    }
    T* dst = p.out + static_cast<int64_t>(row) * p.ldo + f;
    if (live == VEC && (p.ldo % VEC == 0) && (reinterpret_cast<uintptr_t>(p.out) % 16 == 0)) {
        store_vec<VEC>(dst, acc);
    } else {
        for (int e = 0; e < live; ++e) {
            float one[1] = {acc[e]};
            store_vec<1>(dst + e, one);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Narrow fp32 rows staged through SHARED MEMORY (cp.async): the rows in flight live in shared memory, not registers.
//
// The register kernels above hold every in-flight row in registers: NB rows per (half-)warp, ~256 rows per SM at
// full occupancy, and each step of a row's walk waits a whole DRAM round trip.  Here a warp owns 32 work items and
// walks them in BATCHES (<= CAP stored entries of one item, i.e. normally the whole item): it issues the row copies
// of a batch with cp.async (LDGSTS.128: global -> shared, no register, completion tracked per commit group), waits for
// the group, and accumulates the batch out of shared memory; the column indices / values of the next batch are fetched
// one batch ahead.  S > 1 makes it a software pipeline (copies of batch b + S - 1 issued before batch b is consumed);
// measured, ONE stage per warp and twice the resident warps wins (plnlp_spmm_tune mode 10, the default).
// Citation2-shape graph (profiles/r02_spmm_tune_ab_run*.txt, r02_spmm_insitu.txt): F = 50 on its own pitch 3.88 ->
// 2.97 ms (55 % -> 72 % of the HBM copy peak; every 200-byte row costs four 64-byte DRAM sectors, so 78 % is the
// ceiling of the layout), on a 64-float pitch 3.79 -> 2.47 ms, F = 64 3.78 -> 3.0 ms, F = 32 3.40 -> 1.9 ms.
//   issue   : lane l = (sub, part): row j0 + sub of the batch, 16-byte piece `part` of it (LPR pieces per row, so one
//             LDGSTS instruction copies 32 / LPR rows); the values of the batch go to the stage with one STS
//   consume : lane l owns floats 2l, 2l + 1 of the row: one conflict-free LDS.64 per stored entry, strictly in CSR
//             order, separate multiply and add -> the same bits as the register kernels
//   WINDOW  : rows on a pitch of 4k + 2 floats (the 50-wide embedding table on its own pitch) start 8 bytes off a
//             16-byte boundary every other row.  The copy then takes the 16-byte aligned window around the row
//             (8-byte pieces cost an LDGSTS instruction per row and were SLOWER than the register kernel: 4.84 vs
//             3.88 ms); the entry record in the stage tells the consumer where the row's first float landed.
//             The window of the LAST row of x may end past the operand: its last piece is cut to 8 bytes.
// Empty items still produce their (bias / zero) row: every item yields at least one batch.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// predicated forms: the copy is issued iff ok != 0 (no branch around the address arithmetic)
__device__ __forceinline__ void cp_async_16_if(uint32_t dst, const void* src, int ok) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p cp.async.cg.shared.global [%0], [%1], 16;\n}" ::"r"(dst),
                 "l"(src), "r"(ok)
                 : "memory");
}
__device__ __forceinline__ void cp_async_8_if(uint32_t dst, const void* src, int ok) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p cp.async.ca.shared.global [%0], [%1], 8;\n}" ::"r"(dst),
                 "l"(src), "r"(ok)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int STAGED_IPW = 32;      // work items per warp

template <bool WINDOW, bool HAS_VAL, int CAP, int S>
__global__ void __launch_bounds__(512) spmm_csr_staged_kernel(const SpmmParams p, const int lpr, const int rowb,
                                                              const int last_c) {
    static_assert(CAP <= 32 && S >= 1, "one column index per lane and batch");
    constexpr int ENT = WINDOW ? 8 : 4;                // per-entry record: value (+ byte offset of the row in the stage)
    extern __shared__ __align__(16) uint8_t staged_smem[];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // Warp w owns items w, w + W, w + 2 W, ... (W = warps of the grid), NOT 32 consecutive ones: the chunks of a hub
    // row are consecutive items, and 25 chunks of 1024 entries in one warp are 37 x the average warp's work -- that
    // warp alone ran as long as the rest of the kernel (ncu on the bench graph: SMs idle 26 % of the elapsed cycles,
    // 4.15 ms instead of 3.16).  Strided, every warp gets at most one chunk of a given hub.
    const int64_t n_warps = ceil_div(p.n_items, STAGED_IPW);
    const int64_t w = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + wib;
    if (w >= n_warps) return;                          // whole warp; the kernel has no block-wide barrier
    const int cnt = static_cast<int>(min(static_cast<int64_t>(STAGED_IPW), ceil_div(p.n_items - w, n_warps)));
    const int64_t my_item = w + static_cast<int64_t>(lane) * n_warps;

    const int stage_bytes = CAP * rowb + CAP * ENT;    // rows, then the batch's entry records
    uint8_t* const wbase = staged_smem + static_cast<size_t>(wib) * S * stage_bytes;
    const uint32_t wbase_s = static_cast<uint32_t>(__cvta_generic_to_shared(wbase));

    int my_beg = 0, my_end = 0, my_row = 0, my_slot = -1;
    if (lane < cnt) {
        my_beg = __ldg(p.item_ptr + my_item);
        my_end = p.item_end ? __ldg(p.item_end + my_item) : __ldg(p.item_ptr + my_item + 1);
        my_row = __ldg(p.item_row + my_item);
        my_slot = __ldg(p.item_slot + my_item);
    }

    // issue mapping
    const int sub = lane / lpr, part = lane - sub * lpr;
    const int rpi = 32 / lpr;
    const bool issuer = sub < rpi;
    const uint32_t dpart = static_cast<uint32_t>(part * 16);
    const bool tail_lane = part == lpr - 1;
    // consume mapping
    const int f = 2 * lane;
    const bool act = f < p.F;

    // batch iterators (warp-uniform): item i, entries [base, min(base + CAP, end)); the producer runs S - 1 ahead
    int pi = 0, pbase = __shfl_sync(FULL, my_beg, 0), pend = __shfl_sync(FULL, my_end, 0);
    int ci = 0, cbase = pbase, cend = pend;
    int pstage = 0, cstage = 0;

    // column indices / values of the producer's current batch, fetched one batch ahead
    int c_cur = 0;
    float v_cur = 0.0f;
    auto fetch_cols = [&]() {
        const int n = max(0, min(CAP, pend - pbase));
        c_cur = 0;
        v_cur = 0.0f;
        if (lane < n) {
            c_cur = __ldg(p.col + pbase + lane);
            if (HAS_VAL) v_cur = __ldg(p.val + pbase + lane);
        }
    };
    auto issue = [&]() {          // copies of the producer's current batch -> stage pstage; advance; fetch ahead
        const int n = max(0, min(CAP, pend - pbase));
        uint8_t* const stage = wbase + pstage * stage_bytes;
        const uint32_t rows_s = wbase_s + static_cast<uint32_t>(pstage * stage_bytes);
        // element offset of a row inside x fits 32 bits (staged_ok): one IMAD + one IMAD.WIDE per address
        const uint32_t ldx32 = static_cast<uint32_t>(p.ldx);
        const char* const xb = reinterpret_cast<const char*>(p.x);
        bool has_last = false;
        if (WINDOW) {
            // entry record: (value, where the row's first float sits in the stage -- 8 bytes into its slot when the
            // row starts off a 16-byte boundary; x itself is 16-byte aligned, so that is bit 1 of the element offset)
            if (lane < n) {
                const uint32_t eoff = static_cast<uint32_t>(c_cur) * ldx32;
                const int off = lane * rowb + static_cast<int>((eoff & 2u) << 2);
                *reinterpret_cast<float2*>(stage + CAP * rowb + lane * 8) = make_float2(v_cur, __int_as_float(off));
            }
            // the window of an aligned row ends 8 bytes past the row: inside the next row, except for the LAST row
            // of x, whose last piece is cut to the 8 bytes that exist (a batch that holds it takes the careful loop)
            has_last = __any_sync(FULL, lane < n && c_cur == last_c);
        } else if (HAS_VAL) {
            if (lane < n) *reinterpret_cast<float*>(stage + CAP * rowb + lane * 4) = v_cur;
        }
        const uint32_t dst0 = rows_s + dpart + static_cast<uint32_t>(sub * rowb);
        if (!has_last) {
#pragma unroll 4
            for (int j0 = 0; j0 < n; j0 += rpi) {
                const int j = j0 + sub;
                const uint32_t eoff = static_cast<uint32_t>(__shfl_sync(FULL, c_cur, j & 31)) * ldx32;
                const uint32_t e16 = WINDOW ? (eoff & ~3u) : eoff;
                cp_async_16_if(dst0 + static_cast<uint32_t>(j0 * rowb), xb + static_cast<size_t>(e16) * 4 + dpart,
                               issuer && j < n);
            }
        } else {
#pragma unroll 1
            for (int j0 = 0; j0 < n; j0 += rpi) {
                const int j = j0 + sub;
                const int cj = __shfl_sync(FULL, c_cur, j & 31);
                const uint32_t eoff = static_cast<uint32_t>(cj) * ldx32;
                const char* src = xb + static_cast<size_t>(eoff & ~3u) * 4 + dpart;
                const bool cut = tail_lane && cj == last_c && !(eoff & 2u);
                cp_async_16_if(dst0 + static_cast<uint32_t>(j0 * rowb), src, issuer && j < n && !cut);
                cp_async_8_if(dst0 + static_cast<uint32_t>(j0 * rowb), src, issuer && j < n && cut);
            }
        }
        pbase += CAP;
        if (pbase >= pend) {
            ++pi;
            if (pi < cnt) {
                pbase = __shfl_sync(FULL, my_beg, pi);
                pend = __shfl_sync(FULL, my_end, pi);
            }
        }
        if (pi < cnt) fetch_cols();
    };

    fetch_cols();
#pragma unroll 1
    for (int k = 0; k < S - 1; ++k) {
        if (pi < cnt) issue();
        cp_async_commit();
        pstage = (pstage + 1 == S) ? 0 : pstage + 1;
    }

    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 1
    while (ci < cnt) {
        if (pi < cnt) issue();
        cp_async_commit();
        pstage = (pstage + 1 == S) ? 0 : pstage + 1;
        cp_async_wait<S - 1>();                        // this lane's copies of the oldest batch have landed ...
        __syncwarp();                                  // ... and so have every other lane's

        const int n = max(0, min(CAP, cend - cbase));
        const uint8_t* rows = wbase + cstage * stage_bytes;
        const uint8_t* ent = rows + CAP * rowb;
        const uint8_t* mine = rows + 8 * lane;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            float2 t = make_float2(0.0f, 0.0f);
            float v = 1.0f;
            if (WINDOW) {
                const float2 e = *reinterpret_cast<const float2*>(ent + j * 8);
                v = e.x;
                if (act) t = *reinterpret_cast<const float2*>(mine + __float_as_int(e.y));
            } else {
                if (HAS_VAL) v = *reinterpret_cast<const float*>(ent + j * 4);
                if (act) t = *reinterpret_cast<const float2*>(mine + j * rowb);
            }
            if (HAS_VAL) {
                acc0 = __fadd_rn(acc0, __fmul_rn(v, t.x));
                acc1 = __fadd_rn(acc1, __fmul_rn(v, t.y));
            } else {
                acc0 = __fadd_rn(acc0, t.x);
                acc1 = __fadd_rn(acc1, t.y);
            }
        }
        const bool last = cbase + CAP >= cend;
        if (last) {                                    // warp-uniform
            const int row = __shfl_sync(FULL, my_row, ci);
            const int slot = __shfl_sync(FULL, my_slot, ci);
            if (act) {
                float a[2] = {acc0, acc1};
                if (slot >= 0) {
                    store_vec<2>(p.partial + static_cast<int64_t>(slot) * p.F + f, a);
                } else {
                    if (p.row_div) {
                        const float d = __ldg(p.row_div + row);
                        a[0] = a[0] / d; a[1] = a[1] / d;
                    }
                    if (p.bias) { a[0] += __ldg(p.bias + f); a[1] += __ldg(p.bias + f + 1); }
                    if (p.relu) { a[0] = fmaxf(a[0], 0.0f); a[1] = fmaxf(a[1], 0.0f); }
                    if (p.drop_p > 0.0f) {
                        const float sc = 1.0f / (1.0f - p.drop_p);
                        const uint64_t idx = static_cast<uint64_t>(row) * static_cast<uint64_t>(p.F) + f;
                        a[0] = dropout_keep(p.seed, idx, p.drop_p) ? a[0] * sc : 0.0f;
                        a[1] = dropout_keep(p.seed, idx + 1, p.drop_p) ? a[1] * sc : 0.0f;
                    }
                    if (p.mask) {
                        const float* m = p.mask + static_cast<int64_t>(row) * p.ldmask + f;
                        a[0] = __ldg(m) > 0.0f ? a[0] * p.mask_scale : 0.0f;
                        a[1] = __ldg(m + 1) > 0.0f ? a[1] * p.mask_scale : 0.0f;
                    }
                    float* dst = p.out + static_cast<int64_t>(row) * p.ldo + f;
                    if ((p.ldo % 2 == 0) && (reinterpret_cast<uintptr_t>(p.out) % 8 == 0)) {
                        store_vec<2>(dst, a);
                    } else {
                        dst[0] = a[0];
                        dst[1] = a[1];
                    }
                }
            }
            acc0 = acc1 = 0.0f;
        }
        cbase += CAP;
        if (cbase >= cend) {
            ++ci;
            if (ci < cnt) {
                cbase = __shfl_sync(FULL, my_beg, ci);
                cend = __shfl_sync(FULL, my_end, ci);
            }
        }
        cstage = (cstage + 1 == S) ? 0 : cstage + 1;
        __syncwarp();                                  // the stage is free before the next issue overwrites it
    }
}

// second pass for split (hub) rows: sum the partial slots in slot order, then the epilogue
template <typename T, int VEC, int U>
__global__ void __launch_bounds__(256) spmm_fix_kernel(const SpmmParamsT<T> p) {
    const int lane = threadIdx.x & 31;
    const int64_t j = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= p.n_fix) return;
    const int fbase = blockIdx.y * (32 * VEC * U) + lane * VEC;
    const int s0 = __ldg(p.fix_ptr + j), s1 = __ldg(p.fix_ptr + j + 1);
    const int row = __ldg(p.fix_row + j);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int f = fbase + u * 32 * VEC;
        if (f >= p.F) continue;
        float a[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) a[e] = 0.0f;
        for (int s = s0; s < s1; ++s) {
            float t[VEC];
            load_vec<VEC>(t, p.partial + static_cast<int64_t>(s) * p.F + f);
#pragma unroll
            for (int e = 0; e < VEC; ++e) a[e] = __fadd_rn(a[e], t[e]);
        }
        finish_store<T, VEC>(p, row, f, a);
    }
}

template <typename T, int VEC, int U>
static int launch_fix(const SpmmParamsT<T>& p, cudaStream_t st) {
    if (p.n_fix > 0) {
        const int per = 32 * VEC * U;
        const dim3 grid(static_cast<unsigned>(ceil_div(p.n_fix, 8)), static_cast<unsigned>(ceil_div(p.F, per)));
        spmm_fix_kernel<T, VEC, U><<<grid, 256, 0, st>>>(p);
        PLNLP_LAUNCH_CHECK();
    }
    return 0;
}

// narrow rows: see spmm_csr_narrow_kernel.  PLNLP_SPMM_NARROW=0 switches the path off (tuning / A-B runs).
template <typename T>
static bool narrow_ok(const SpmmParamsT<T>& p) {
    static const bool on = [] { const char* e = getenv("PLNLP_SPMM_NARROW"); return !(e && e[0] == '0'); }();
    constexpr int VEC = 16 / sizeof(T);
    return on && p.F <= 16 * VEC && p.x_index == nullptr && (p.ldx % VEC == 0) && p.ldx >= ((p.F + VEC - 1) / VEC) * VEC &&
           aligned(p.x, 16);
}

template <typename T>
static int launch_narrow(const SpmmParamsT<T>& p, cudaStream_t st) {
    static const int nb_env = [] { const char* e = getenv("PLNLP_SPMM_NB"); return e ? atoi(e) : 0; }();
    // two neighbours per half in flight: measured best (F = 64: 3.78 ms vs 3.86 at four; eight spills)
    const int nb = nb_env ? nb_env : 2;
    const dim3 grid(static_cast<unsigned>(ceil_div(ceil_div(p.n_items, 2), 8)));
#define PLNLP_NARROW_LAUNCH(NBV)                                                          \
    do {                                                                                  \
        if (p.val) spmm_csr_narrow_kernel<T, true, NBV><<<grid, 256, 0, st>>>(p);         \
        else       spmm_csr_narrow_kernel<T, false, NBV><<<grid, 256, 0, st>>>(p);        \
    } while (0)
    if (nb >= 4) PLNLP_NARROW_LAUNCH(4);
    else PLNLP_NARROW_LAUNCH(2);
#undef PLNLP_NARROW_LAUNCH
    PLNLP_LAUNCH_CHECK();
    return 0;
}

// shared-memory staged kernel: fp32 rows of an even width <= 64 floats, 16-byte aligned base, no x_index; rows on a
// pitch of 4k floats are copied as they are (a width that is not a multiple of 4 reads pitch padding that is never
// consumed), rows on a pitch of 4k + 2 floats through their aligned window
struct StagedGeom {
    bool window;
    int lpr, rowb, last_c;
};
static bool staged_ok(const SpmmParams& p, int64_t x_rows, StagedGeom& g) {
    if (g_spmm_staged <= 0 || p.F > 64 || (p.F % 2) || p.x_index || (p.ldx % 2) || !aligned(p.x, 16) || x_rows <= 0)
        return false;
    if (p.partial && !aligned(p.partial, 8)) return false;
    g.window = (p.ldx % 4) != 0;
    g.lpr = static_cast<int>(ceil_div(static_cast<int64_t>(p.F) * 4 + (g.window ? 8 : 0), 16));
    g.rowb = g.lpr * 16;
    g.last_c = static_cast<int>(x_rows - 1);
    return g.lpr <= 32 && x_rows < (1ll << 31) && x_rows * p.ldx < (1ll << 32);
}

template <bool WINDOW, bool HAS_VAL, int CAP, int S>
static int launch_staged_one(const SpmmParams& p, const StagedGeom& g, cudaStream_t st) {
    const int64_t n_warps = ceil_div(p.n_items, STAGED_IPW);
    const int stage_bytes = CAP * g.rowb + CAP * (WINDOW ? 8 : 4);
    int warps = g_spmm_staged_warps;
    warps = warps < 1 ? 1 : (warps > 16 ? 16 : warps);
    while (warps > 1 && static_cast<int64_t>(warps) * S * stage_bytes > 200 * 1024) --warps;
    const int bytes = warps * S * stage_bytes;
    auto kern = spmm_csr_staged_kernel<WINDOW, HAS_VAL, CAP, S>;
    static int configured = 0;
    if (bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        // the kernel lives on resident CTAs x shared memory: ask for the full carve-out.  Left to the driver's
        // heuristic the SM kept the L1-heavy split of the neighbouring kernels of a training step and fewer CTAs fitted
        // (same kernel, same data: 2.8 / 3.3 ms back to back in a microbenchmark, 3.6 / 4.2 ms inside the step)
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured = bytes;
    }
    const dim3 grid(static_cast<unsigned>(ceil_div(n_warps, warps)));
    kern<<<grid, warps * 32, bytes, st>>>(p, g.lpr, g.rowb, g.last_c);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

template <bool WINDOW, bool HAS_VAL>
static int launch_staged_cfg(const SpmmParams& p, const StagedGeom& g, cudaStream_t st) {
    int mode = g_spmm_staged;
    if (mode >= 12) {
        // auto: whole items per batch (CAP = 32) and ONE stage per warp -- twice the resident warps of a two-stage
        // ring, each with its whole batch in flight (profiles/r02_spmm_tune_ab.txt, citation2-shape graph, after the
        // strided item assignment: F = 50 on its own pitch 3.00 ms vs 3.17 - 3.47 with two stages, on a 64-float
        // pitch 2.60 vs 2.95 - 3.10)
        mode = 10;
    }
    switch (mode) {                     // (CAP, S): rows per batch, stages per warp
        case 1: return launch_staged_one<WINDOW, HAS_VAL, 16, 4>(p, g, st);
        case 2: return launch_staged_one<WINDOW, HAS_VAL, 32, 3>(p, g, st);
        case 3: return launch_staged_one<WINDOW, HAS_VAL, 16, 6>(p, g, st);
        case 4: return launch_staged_one<WINDOW, HAS_VAL, 32, 4>(p, g, st);
        case 6: return launch_staged_one<WINDOW, HAS_VAL, 16, 2>(p, g, st);
        case 7: return launch_staged_one<WINDOW, HAS_VAL, 8, 3>(p, g, st);
        case 8: return launch_staged_one<WINDOW, HAS_VAL, 8, 4>(p, g, st);
        case 9: return launch_staged_one<WINDOW, HAS_VAL, 32, 2>(p, g, st);
        case 10: return launch_staged_one<WINDOW, HAS_VAL, 32, 1>(p, g, st);
        case 11: return launch_staged_one<WINDOW, HAS_VAL, 16, 1>(p, g, st);
        default: return launch_staged_one<WINDOW, HAS_VAL, 16, 3>(p, g, st);      // 5
    }
}

static int launch_staged(const SpmmParams& p, const StagedGeom& g, cudaStream_t st) {
    if (g.window) return p.val ? launch_staged_cfg<true, true>(p, g, st) : launch_staged_cfg<true, false>(p, g, st);
    return p.val ? launch_staged_cfg<false, true>(p, g, st) : launch_staged_cfg<false, false>(p, g, st);
}

template <typename T, int VEC, int U>
static int launch_spmm(const SpmmParamsT<T>& p, cudaStream_t st) {
    const int per = 32 * VEC * U;
    const unsigned slabs = static_cast<unsigned>(ceil_div(p.F, per));
    const dim3 block(256);
    // loads in flight per lane: NB neighbours x U vectors.  Measured on the citation2-shape graph
    // (tools/spmm_sweep.py, profiles/r01_spmm_sweep.txt): with 16-byte loads two neighbours per lane are enough
    // and leave the most warps resident (F=128: 83 %, F=256: 91 % of the HBM peak; 66 % / 87 % at eight);
    // 8- and 4-byte loads (odd widths such as F=50, bf16 below 256) want four.  PLNLP_SPMM_NB (2, 4 or 8)
    // overrides NB for tuning.
    static const int nb_env = [] { const char* e = getenv("PLNLP_SPMM_NB"); return e ? atoi(e) : 0; }();
    const int nb = nb_env ? nb_env : (VEC * static_cast<int>(sizeof(T)) >= 16 ? 2 : 4);
#define PLNLP_SPMM_LAUNCH(NBV)                                                          \
    do {                                                                                \
        if (p.val) spmm_csr_kernel<T, VEC, U, NBV, true><<<grid, block, 0, st>>>(p);    \
        else       spmm_csr_kernel<T, VEC, U, NBV, false><<<grid, block, 0, st>>>(p);   \
    } while (0)
    if (p.n_items > 0) {
        const dim3 grid(static_cast<unsigned>(ceil_div(p.n_items, 8)), slabs);
        if (nb >= 8) PLNLP_SPMM_LAUNCH(8);
        else if (nb >= 4) PLNLP_SPMM_LAUNCH(4);
        else PLNLP_SPMM_LAUNCH(2);
        PLNLP_LAUNCH_CHECK();
    }
#undef PLNLP_SPMM_LAUNCH
    return launch_fix<T, VEC, U>(p, st);
}

template <typename T, int VEC>
static int dispatch_u(const SpmmParamsT<T>& p, cudaStream_t st) {
    const int64_t lanes_needed = ceil_div(p.F, VEC);
    if (lanes_needed <= 32) return launch_spmm<T, VEC, 1>(p, st);
    if (lanes_needed <= 64 || VEC == 8) return launch_spmm<T, VEC, 2>(p, st);
    return launch_spmm<T, VEC, 4>(p, st);
}

}  // namespace plnlp

extern "C" int plnlp_spmm_tune(int prefetch_mode, int staged_mode, int staged_warps, int l2_fetch_bytes) {
    using namespace plnlp;
    PLNLP_REQUIRE(prefetch_mode >= -1 && prefetch_mode <= 3 && staged_mode >= -1 && staged_mode <= 12, PLNLP_E_SIZE);
    if (prefetch_mode >= 0) g_spmm_pf = prefetch_mode;
    if (staged_mode >= 0) g_spmm_staged = staged_mode;
    if (staged_warps > 0) g_spmm_staged_warps = staged_warps;
    if (l2_fetch_bytes > 0) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, static_cast<size_t>(l2_fetch_bytes));
        if (e != cudaSuccess) {
            cudaGetLastError();
            return static_cast<int>(e);
        }
    }
    return 0;
}

extern "C" int plnlp_spmm_csr_f32(const int32_t* item_ptr, const int32_t* item_row, const int32_t* item_slot,
                                  int64_t n_items, const int32_t* item_end, const int32_t* x_index, const int32_t* col, const float* val,
                                  const float* row_div, const float* bias, int relu, float drop_p,
                                  uint64_t seed, const float* x, int64_t ldx, int64_t x_rows, float* out, int64_t ldo,
                                  int64_t F, float* partial, const int32_t* fix_ptr, const int32_t* fix_row,
                                  int64_t n_fix, const float* mask, int64_t ldmask, float mask_scale, void* stream) {
    using namespace plnlp;
    PLNLP_REQUIRE(n_items >= 0 && n_fix >= 0 && F > 0 && F < (1 << 30), PLNLP_E_SIZE);
    if (n_items == 0) return 0;
    PLNLP_REQUIRE(item_ptr && item_row && item_slot && x && out, PLNLP_E_NULL);
    PLNLP_REQUIRE(ldx >= F && ldo >= F, PLNLP_E_SIZE);
    PLNLP_REQUIRE(drop_p >= 0.0f && drop_p < 1.0f, PLNLP_E_SIZE);
    if (n_fix > 0) PLNLP_REQUIRE(partial && fix_ptr && fix_row, PLNLP_E_NULL);
    SpmmParams p{item_ptr, item_row, item_slot, n_items, item_end, x_index, col, val, row_div, bias, relu, drop_p, seed,
                 x, ldx, out, ldo, static_cast<int>(F), partial, fix_ptr, fix_row, n_fix, mask, ldmask, mask_scale,
                 resolve_pf(F * 4, 4)};
    PLNLP_REQUIRE(!mask || ldmask >= F, PLNLP_E_SIZE);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool v4 = (F % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && aligned(x, 16) && aligned(out, 16) &&
                    (!partial || aligned(partial, 16));
    const bool v2 = (F % 2 == 0) && (ldx % 2 == 0) && (ldo % 2 == 0) && aligned(x, 8) && aligned(out, 8) &&
                    (!partial || aligned(partial, 8));
    StagedGeom sg;
    const bool staged = staged_ok(p, x_rows, sg);
    if (staged || narrow_ok(p)) {
        int rc = staged ? launch_staged(p, sg, st) : launch_narrow(p, st);
        if (rc != 0) return rc;
        // hub rows: the fixed-order combine of the partial slots (any vector width the partial buffer allows)
        const bool f4 = (F % 4 == 0) && (ldo % 4 == 0) && aligned(out, 16) && (!partial || aligned(partial, 16));
        const bool f2 = (F % 2 == 0) && (ldo % 2 == 0) && aligned(out, 8) && (!partial || aligned(partial, 8));
        if (f4) return launch_fix<float, 4, 1>(p, st);
        if (f2) return launch_fix<float, 2, 1>(p, st);
        return launch_fix<float, 1, 2>(p, st);
    }
    if (v4) return dispatch_u<float, 4>(p, st);
    if (v2) return dispatch_u<float, 2>(p, st);
    return dispatch_u<float, 1>(p, st);
}

// bf16 feature storage (x and out are bf16 bit patterns), fp32 accumulation in CSR order, one RN rounding at
// the store.  Same plan, epilogue and Philox indexing as the fp32 entry point; `partial` stays fp32.
extern "C" int plnlp_spmm_csr_bf16(const int32_t* item_ptr, const int32_t* item_row, const int32_t* item_slot,
                                   int64_t n_items, const int32_t* item_end, const int32_t* x_index, const int32_t* col, const float* val,
                                   const float* row_div, const float* bias, int relu, float drop_p,
                                   uint64_t seed, const uint16_t* x, int64_t ldx, int64_t x_rows, uint16_t* out, int64_t ldo,
                                   int64_t F, float* partial, const int32_t* fix_ptr, const int32_t* fix_row,
                                   int64_t n_fix, const float* mask, int64_t ldmask, float mask_scale, void* stream) {
    using namespace plnlp;
    PLNLP_REQUIRE(n_items >= 0 && n_fix >= 0 && F > 0 && F < (1 << 30), PLNLP_E_SIZE);
    if (n_items == 0) return 0;
    PLNLP_REQUIRE(item_ptr && item_row && item_slot && x && out, PLNLP_E_NULL);
    PLNLP_REQUIRE(ldx >= F && ldo >= F, PLNLP_E_SIZE);
    PLNLP_REQUIRE(drop_p >= 0.0f && drop_p < 1.0f, PLNLP_E_SIZE);
    if (n_fix > 0) PLNLP_REQUIRE(partial && fix_ptr && fix_row, PLNLP_E_NULL);
    SpmmParamsT<__nv_bfloat16> p{item_ptr, item_row, item_slot, n_items, item_end, x_index, col, val, row_div, bias, relu, drop_p, seed,
                                 reinterpret_cast<const __nv_bfloat16*>(x), ldx,
                                 reinterpret_cast<__nv_bfloat16*>(out), ldo, static_cast<int>(F), partial, fix_ptr,
                                 fix_row, n_fix, mask, ldmask, mask_scale, resolve_pf(F * 2, 2)};
    PLNLP_REQUIRE(!mask || ldmask >= F, PLNLP_E_SIZE);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto ok = [&](int v) {   // v bf16 elements per access: 2v bytes for x / out, 4v (<= 16-byte pieces) for partial
        return (F % v == 0) && (ldx % v == 0) && (ldo % v == 0) && aligned(x, 2 * v) && aligned(out, 2 * v) &&
               (!partial || aligned(partial, v >= 4 ? 16 : 4 * v));
    };
    // 16-byte loads as soon as a row is longer than 256 bytes (F = 200: 25 lanes of one warp; 9.54 -> 8.17 ms on the
    // citation2-shape graph), else the widest access that still keeps a full warp busy on one row (F / v >= 32), else
    // the widest legal one.  (The two-rows-per-warp kernel was tried for bf16 rows of <= 256 bytes and was SLOWER than
    // this path -- F = 64: 4.88 vs 3.67 ms -- the unpacking of 8 elements per load costs more than the extra rows in
    // flight gain; it stays an fp32 kernel.)
    if (ok(8) && F > 128) return dispatch_u<__nv_bfloat16, 8>(p, st);
    if (ok(4) && F >= 128) return dispatch_u<__nv_bfloat16, 4>(p, st);
    if (ok(2)) return dispatch_u<__nv_bfloat16, 2>(p, st);
    if (ok(4)) return dispatch_u<__nv_bfloat16, 4>(p, st);
    return dispatch_u<__nv_bfloat16, 1>(p, st);
}
