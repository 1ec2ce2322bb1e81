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
};
using SpmmParams = SpmmParamsT<float>;

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
            unsigned live = __ballot_sync(0xffffffffu, lane < n && c >= 0);
#pragma unroll 1
            while (live) gather_block_masked<T, VEC, U, NB, HAS_VAL>(xb, p.ldx, c, v, live, act, acc);
        } else if (n == 32) {
#pragma unroll 1
            for (int j = 0; j < 32; j += NB)
                gather_block<T, VEC, U, NB, HAS_VAL, false>(xb, p.ldx, c, v, j, n, act, acc);
        } else {
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

extern "C" int plnlp_spmm_csr_f32(const int32_t* item_ptr, const int32_t* item_row, const int32_t* item_slot,
                                  int64_t n_items, const int32_t* item_end, const int32_t* x_index, const int32_t* col, const float* val,
                                  const float* row_div, const float* bias, int relu, float drop_p,
                                  uint64_t seed, const float* x, int64_t ldx, float* out, int64_t ldo,
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
                 x, ldx, out, ldo, static_cast<int>(F), partial, fix_ptr, fix_row, n_fix, mask, ldmask, mask_scale};
    PLNLP_REQUIRE(!mask || ldmask >= F, PLNLP_E_SIZE);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool v4 = (F % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && aligned(x, 16) && aligned(out, 16) &&
                    (!partial || aligned(partial, 16));
    const bool v2 = (F % 2 == 0) && (ldx % 2 == 0) && (ldo % 2 == 0) && aligned(x, 8) && aligned(out, 8) &&
                    (!partial || aligned(partial, 8));
    if (narrow_ok(p)) {
        int rc = launch_narrow(p, st);
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
                                   uint64_t seed, const uint16_t* x, int64_t ldx, uint16_t* out, int64_t ldo,
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
                                 fix_row, n_fix, mask, ldmask, mask_scale};
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
