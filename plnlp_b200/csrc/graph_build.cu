// CSR construction on the GPU: the index work main.py does once per run before the hot path starts
// (/root/reference/main.py:81-83 ToSparseTensor, :109-110 to_symmetric, plnlp/utils.py:83-89 set_diag + D^-1/2 A D^-1/2),
// bit-exact in every index array with the torch_sparse semantics restated in oracle/sparse.py:
//
//   * an entry is the 64-bit key row * n_cols + col; "sort by (row, col), duplicates kept, original order among equal
//     keys" is one STABLE radix sort of (key, position) pairs (cub::DeviceRadixSort over the key bits actually used);
//   * to_symmetric = both orientations of every key, sorted, equal keys merged (values summed in sorted order);
//   * set_diag = drop the stored diagonal, insert one unit entry per row -- done in key space, so the result comes
//     out of the same sort;
//   * keys -> (rowptr, col): every thread looks at one sorted key and its predecessor and fills the row pointers of
//     the rows that start between them (no atomics, no histogram);
//   * gcn_normalization values: val[e] = dis[row] * v[e] * dis[col], dis = deg^-1/2 with inf -> 0, deg = row sum of
//     the values (or the row count), left to right like torch.
//
// The sorts are CUB (a library primitive, like torch.sort was); everything around them is written here.  All buffers
// are caller-owned; the element counts that depend on the data (unique keys, kept entries) are returned through a
// device counter the caller reads once.
#include <cub/cub.cuh>

#include "common.cuh"

namespace plnlp {
namespace {

__global__ void __launch_bounds__(256) make_keys_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                                        int64_t n, int64_t n_cols, int both, int drop_diag,
                                                        int64_t* __restrict__ keys, int64_t* __restrict__ pos) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = __ldg(row + i), c = __ldg(col + i);
    // a dropped entry gets the largest key: it sorts to the end and the caller cuts it off
    const int64_t dead = INT64_MAX;
    keys[i] = (drop_diag && r == c) ? dead : r * n_cols + c;
    if (pos) pos[i] = i;
    if (both) {
        keys[n + i] = (drop_diag && r == c) ? dead : c * n_cols + r;
        if (pos) pos[n + i] = n + i;
    }
}

__global__ void __launch_bounds__(256) diag_keys_kernel(int64_t n_diag, int64_t n_cols, int64_t* __restrict__ keys,
                                                        int64_t* __restrict__ pos, int64_t pos0) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_diag) return;
    keys[i] = i * n_cols + i;
    if (pos) pos[i] = pos0 + i;
}

// sorted keys -> rowptr (int64 [n_rows + 1]) and col (int64 [n])
__global__ void __launch_bounds__(256) keys_to_csr_kernel(const int64_t* __restrict__ keys, int64_t n, int64_t n_rows,
                                                          int64_t n_cols, int64_t* __restrict__ rowptr,
                                                          int64_t* __restrict__ col) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i > n) return;
    // rows r with prev_row < r <= this_row start at entry i (i == n: the rows after the last entry, and rowptr[n_rows])
    const int64_t prev_row = i == 0 ? -1 : keys[i - 1] / n_cols;
    const int64_t this_row = i == n ? n_rows : keys[i] / n_cols;
    for (int64_t r = prev_row + 1; r <= this_row; ++r) rowptr[r] = i;
    if (i < n) col[i] = keys[i] - this_row * n_cols;
}

// segment heads of a sorted key array: head[i] = 1 where keys[i] != keys[i-1]
__global__ void __launch_bounds__(256) heads_kernel(const int64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ head) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// merged values of equal keys: out[seg] = sum of val[pos[i]] over the entries of segment seg, in sorted order
__global__ void __launch_bounds__(256) merge_values_kernel(const int32_t* __restrict__ seg_of /* inclusive scan of heads */,
                                                           const int32_t* __restrict__ head, const int64_t* __restrict__ pos,
                                                           const float* __restrict__ val, int64_t n_src, int64_t n,
                                                           float* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    float acc = 0.0f;
    int64_t j = i;
    do {
        const int64_t p = pos[j];
        acc += val[p < n_src ? p : p - n_src];          // the second orientation of entry p - n_src carries its value
        ++j;
    } while (j < n && !head[j]);
    out[seg_of[i] - 1] = acc;
}

__global__ void __launch_bounds__(256) row_sum_kernel(const int64_t* __restrict__ rowptr, const float* __restrict__ val,
                                                      int64_t n_rows, float* __restrict__ dis) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int64_t b = rowptr[r], e = rowptr[r + 1];
    float deg;
    if (val) {
        deg = 0.0f;
        for (int64_t p = b; p < e; ++p) deg += val[p];
    } else {
        deg = static_cast<float>(e - b);
    }
    const float d = 1.0f / sqrtf(deg);                   // deg^-1/2; inf (deg = 0) -> 0 (utils.py:86-87)
    dis[r] = isinf(d) ? 0.0f : d;
}

__global__ void __launch_bounds__(256) sym_norm_kernel(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ col,
                                                       const float* __restrict__ val_in, const float* __restrict__ dis,
                                                       int64_t n_rows, float* __restrict__ val_out) {
    // one warp per row
    const int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    const float dr = dis[r];
    for (int64_t p = rowptr[r] + lane; p < rowptr[r + 1]; p += 32) {
        const float v = val_in ? val_in[p] : 1.0f;
        val_out[p] = __fmul_rn(__fmul_rn(dr, v), dis[col[p]]);      // (dis[:, None] * A) * dis[None, :]
    }
}

int bits_for(int64_t n_rows, int64_t n_cols) {
    // the largest key is n_rows * n_cols - 1; dead keys (INT64_MAX) need every bit
    int b = 1;
    const unsigned __int128 top = static_cast<unsigned __int128>(n_rows) * static_cast<unsigned __int128>(n_cols);
    while (b < 63 && (static_cast<unsigned __int128>(1) << b) < top) ++b;
    return b;
}

}  // namespace
}  // namespace plnlp

using namespace plnlp;

// ---- keys -------------------------------------------------------------------------------------------------------
// keys[i] = row[i] * n_cols + col[i] (and, when `both`, keys[n + i] = col[i] * n_cols + row[i]); pos = 0 .. (2)n - 1
// (may be NULL); drop_diag: entries with row == col get the key INT64_MAX (they sort behind everything else)
extern "C" int plnlp_graph_make_keys(const int64_t* row, const int64_t* col, int64_t n, int64_t n_cols, int both,
                                     int drop_diag, int64_t* keys, int64_t* pos, void* stream) {
    PLNLP_REQUIRE(n >= 0 && n_cols > 0, PLNLP_E_SIZE);
    if (n == 0) return 0;
    PLNLP_REQUIRE(row && col && keys, PLNLP_E_NULL);
    make_keys_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        row, col, n, n_cols, both, drop_diag, keys, pos);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

// keys[i] = i * n_cols + i for i < n_diag (the unit diagonal of set_diag), pos[i] = pos0 + i
extern "C" int plnlp_graph_diag_keys(int64_t n_diag, int64_t n_cols, int64_t* keys, int64_t* pos, int64_t pos0,
                                     void* stream) {
    PLNLP_REQUIRE(n_diag >= 0 && n_cols > 0, PLNLP_E_SIZE);
    if (n_diag == 0) return 0;
    PLNLP_REQUIRE(keys, PLNLP_E_NULL);
    diag_keys_kernel<<<static_cast<unsigned>(ceil_div(n_diag, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n_diag, n_cols, keys, pos, pos0);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

// ---- stable sort of (key, pos) pairs ---------------------------------------------------------------------------------
extern "C" int64_t plnlp_graph_sort_workspace_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const int64_t*>(nullptr), static_cast<int64_t*>(nullptr),
                                    static_cast<const int64_t*>(nullptr), static_cast<int64_t*>(nullptr),
                                    static_cast<int64_t>(n > 0 ? n : 1), 0, 64, static_cast<cudaStream_t>(nullptr));
    return static_cast<int64_t>(bytes) + 256;
}

// keys_out / pos_out = (keys_in, pos_in) sorted by key, equal keys in their input order (pos may be NULL: keys only).
// has_dead: some keys are INT64_MAX (all 63 value bits are then sorted), otherwise only the bits n_rows * n_cols needs.
extern "C" int plnlp_graph_sort_pairs(const int64_t* keys_in, const int64_t* pos_in, int64_t n, int64_t n_rows,
                                      int64_t n_cols, int has_dead, int64_t* keys_out, int64_t* pos_out, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
    PLNLP_REQUIRE(n >= 0 && n < (int64_t(1) << 31), PLNLP_E_SIZE);
    if (n == 0) return 0;
    PLNLP_REQUIRE(keys_in && keys_out && workspace, PLNLP_E_NULL);
    PLNLP_REQUIRE(workspace_bytes >= plnlp_graph_sort_workspace_bytes(n), PLNLP_E_WORKSPACE);
    size_t bytes = static_cast<size_t>(workspace_bytes);
    const int end_bit = has_dead ? 63 : bits_for(n_rows, n_cols);
    cudaError_t e;
    if (pos_in && pos_out)
        e = cub::DeviceRadixSort::SortPairs(workspace, bytes, keys_in, keys_out, pos_in, pos_out, n, 0, end_bit,
                                            static_cast<cudaStream_t>(stream));
    else
        e = cub::DeviceRadixSort::SortKeys(workspace, bytes, keys_in, keys_out, n, 0, end_bit,
                                           static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return static_cast<int>(e);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

// ---- merge equal keys of a sorted array --------------------------------------------------------------------------------
extern "C" int64_t plnlp_graph_unique_workspace_bytes(int64_t n) {
    size_t a = 0, b = 0;
    const int64_t m = n > 0 ? n : 1;
    cub::DeviceSelect::Unique(nullptr, a, static_cast<const int64_t*>(nullptr), static_cast<int64_t*>(nullptr),
                              static_cast<int64_t*>(nullptr), m, static_cast<cudaStream_t>(nullptr));
    cub::DeviceScan::InclusiveSum(nullptr, b, static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr), m,
                                  static_cast<cudaStream_t>(nullptr));
    return static_cast<int64_t>(a > b ? a : b) + 2 * m * 4 + 512;
}

// keys_out = the distinct keys of the SORTED keys_in (at most n), *n_out (device) = how many.  With val != NULL also
// val_out[s] = sum of val over the entries of segment s in sorted order, where entry i took its value from
// val[pos[i] mod n_src] (pos as produced by plnlp_graph_make_keys(both = 1) + plnlp_graph_sort_pairs).
extern "C" int plnlp_graph_unique(const int64_t* keys_in, int64_t n, int64_t* keys_out, int64_t* n_out, const int64_t* pos,
                                  const float* val, int64_t n_src, float* val_out, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
    PLNLP_REQUIRE(n >= 0 && n < (int64_t(1) << 31), PLNLP_E_SIZE);
    PLNLP_REQUIRE(n_out, PLNLP_E_NULL);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        cudaMemsetAsync(n_out, 0, 8, st);
        return 0;
    }
    PLNLP_REQUIRE(keys_in && keys_out && workspace, PLNLP_E_NULL);
    PLNLP_REQUIRE(workspace_bytes >= plnlp_graph_unique_workspace_bytes(n), PLNLP_E_WORKSPACE);
    PLNLP_REQUIRE(!val || (pos && val_out && n_src > 0), PLNLP_E_NULL);
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    int32_t* head = reinterpret_cast<int32_t*>(base);
    int32_t* seg = head + n;
    void* cub_ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(seg + n) + 255) & ~uintptr_t(255));
    size_t cub_bytes = static_cast<size_t>(workspace_bytes) - (reinterpret_cast<uint8_t*>(cub_ws) - reinterpret_cast<uint8_t*>(workspace));
    cudaError_t e = cub::DeviceSelect::Unique(cub_ws, cub_bytes, keys_in, keys_out, n_out, n, st);
    if (e != cudaSuccess) return static_cast<int>(e);
    PLNLP_LAUNCH_CHECK();
    if (val) {
        heads_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, st>>>(keys_in, n, head);
        PLNLP_LAUNCH_CHECK();
        e = cub::DeviceScan::InclusiveSum(cub_ws, cub_bytes, head, seg, n, st);
        if (e != cudaSuccess) return static_cast<int>(e);
        PLNLP_LAUNCH_CHECK();
        merge_values_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, st>>>(seg, head, pos, val, n_src, n, val_out);
        PLNLP_LAUNCH_CHECK();
    }
    return 0;
}

// ---- sorted keys -> CSR ---------------------------------------------------------------------------------------------
extern "C" int plnlp_graph_keys_to_csr(const int64_t* keys, int64_t n, int64_t n_rows, int64_t n_cols, int64_t* rowptr,
                                       int64_t* col, void* stream) {
    PLNLP_REQUIRE(n >= 0 && n_rows >= 0 && n_cols > 0, PLNLP_E_SIZE);
    PLNLP_REQUIRE(rowptr && (n == 0 || (keys && col)), PLNLP_E_NULL);
    keys_to_csr_kernel<<<static_cast<unsigned>(ceil_div(n + 1, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        keys, n, n_rows, n_cols, rowptr, col);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

// ---- D^-1/2 A D^-1/2 (utils.py:83-89 after set_diag) ---------------------------------------------------------------------
// dis [n_rows] (scratch / output): deg^-1/2 with inf -> 0; val_out[e] = dis[row] * val_in[e] * dis[col] (val_in NULL: 1)
extern "C" int plnlp_graph_sym_normalize(const int64_t* rowptr, const int64_t* col, const float* val_in, int64_t n_rows,
                                         float* dis, float* val_out, void* stream) {
    PLNLP_REQUIRE(n_rows >= 0, PLNLP_E_SIZE);
    if (n_rows == 0) return 0;
    PLNLP_REQUIRE(rowptr && col && dis && val_out, PLNLP_E_NULL);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    row_sum_kernel<<<static_cast<unsigned>(ceil_div(n_rows, 256)), 256, 0, st>>>(rowptr, val_in, n_rows, dis);
    PLNLP_LAUNCH_CHECK();
    sym_norm_kernel<<<static_cast<unsigned>(ceil_div(n_rows * 32, 256)), 256, 0, st>>>(rowptr, col, val_in, dis, n_rows, val_out);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

// ---- row-subset SpMM plan (plnlp_b200/graph.py build_subset_plan) -----------------------------------------------------
// The last conv computes only the rows the batch reads (model.py:152-156): per step, a plan for T selected rows of
// the adjacency.  Two kernels around three prefix sums instead of ~25 torch index ops and three host reads:
//   count: n_it[t] = max(1, ceil(len_t / chunk)) items for row rows[t]; multi[t] = n_it[t] > 1
//   (caller: exclusive prefix sums of n_it -> first, of multi -> fix index, of multi ? n_it : 0 -> partial slot base)
//   fill : item arrays, fix arrays, mean divisors
namespace plnlp {
namespace {

__global__ void __launch_bounds__(256) subset_count_kernel(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ rows,
                                                           int64_t T, int chunk, int64_t* __restrict__ n_it,
                                                           int64_t* __restrict__ multi, int64_t* __restrict__ n_slot,
                                                           int64_t* __restrict__ len_out) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int64_t r = rows[t];
    const int64_t len = rowptr[r + 1] - rowptr[r];
    int64_t n = (len + chunk - 1) / chunk;
    if (n < 1) n = 1;
    n_it[t] = n;
    multi[t] = n > 1 ? 1 : 0;
    n_slot[t] = n > 1 ? n : 0;
    len_out[t] = len;
}

__global__ void __launch_bounds__(256) subset_fill_kernel(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ rows,
                                                          int64_t T, int chunk, const int64_t* __restrict__ first /* inclusive */,
                                                          const int64_t* __restrict__ fixi /* inclusive */,
                                                          const int64_t* __restrict__ slot /* inclusive */,
                                                          const int64_t* __restrict__ n_it, int32_t* __restrict__ item_ptr,
                                                          int32_t* __restrict__ item_end, int32_t* __restrict__ item_row,
                                                          int32_t* __restrict__ item_slot, int32_t* __restrict__ fix_ptr,
                                                          int32_t* __restrict__ fix_row, float* __restrict__ row_cnt) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int64_t r = rows[t];
    const int64_t beg = rowptr[r], end = rowptr[r + 1];
    const int64_t n = n_it[t];
    const int64_t i0 = first[t] - n;                       // exclusive prefix
    const bool multi = n > 1;
    const int64_t s0 = multi ? slot[t] - n : -1;
    for (int64_t k = 0; k < n; ++k) {
        const int64_t b = beg + k * chunk;
        const int64_t e = b + chunk < end ? b + chunk : end;
        item_ptr[i0 + k] = static_cast<int32_t>(b);
        item_end[i0 + k] = static_cast<int32_t>(e);
        item_row[i0 + k] = static_cast<int32_t>(t);
        item_slot[i0 + k] = multi ? static_cast<int32_t>(s0 + k) : -1;
    }
    if (multi) {
        const int64_t j = fixi[t] - 1;
        fix_row[j] = static_cast<int32_t>(t);
        fix_ptr[j] = static_cast<int32_t>(s0);
        // the closing entry fix_ptr[n_fix] = n_partial is written by whoever owns the last multi row
        fix_ptr[j + 1] = static_cast<int32_t>(s0 + n);
    }
    const int64_t len = end - beg;
    row_cnt[t] = static_cast<float>(len > 1 ? len : 1);
}

}  // namespace
}  // namespace plnlp

extern "C" int plnlp_subset_plan_count(const int64_t* rowptr, const int64_t* rows, int64_t T, int chunk, int64_t* n_it,
                                       int64_t* multi, int64_t* n_slot, int64_t* len, void* stream) {
    PLNLP_REQUIRE(T >= 0 && chunk > 0, PLNLP_E_SIZE);
    if (T == 0) return 0;
    PLNLP_REQUIRE(rowptr && rows && n_it && multi && n_slot && len, PLNLP_E_NULL);
    plnlp::subset_count_kernel<<<static_cast<unsigned>(plnlp::ceil_div(T, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        rowptr, rows, T, chunk, n_it, multi, n_slot, len);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

// first / fixi / slot: INCLUSIVE prefix sums of n_it / multi / n_slot.  fix_ptr has n_fix + 1 entries (at least 1).
extern "C" int plnlp_subset_plan_fill(const int64_t* rowptr, const int64_t* rows, int64_t T, int chunk, const int64_t* first,
                                      const int64_t* fixi, const int64_t* slot, const int64_t* n_it, int32_t* item_ptr,
                                      int32_t* item_end, int32_t* item_row, int32_t* item_slot, int32_t* fix_ptr,
                                      int32_t* fix_row, float* row_cnt, void* stream) {
    PLNLP_REQUIRE(T >= 0 && chunk > 0, PLNLP_E_SIZE);
    if (T == 0) return 0;
    PLNLP_REQUIRE(rowptr && rows && first && fixi && slot && n_it && item_ptr && item_end && item_row && item_slot &&
                      fix_ptr && row_cnt, PLNLP_E_NULL);
    plnlp::subset_fill_kernel<<<static_cast<unsigned>(plnlp::ceil_div(T, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        rowptr, rows, T, chunk, first, fixi, slot, n_it, item_ptr, item_end, item_row, item_slot, fix_ptr, fix_row, row_cnt);
    PLNLP_LAUNCH_CHECK();
    return 0;
}
