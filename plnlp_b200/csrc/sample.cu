// Negative samplers on the GPU.
//
// Replaces /root/reference/plnlp/negative_sample.py:31-43 (local_neg_sample: CPU torch.randint +
// H2D copy of [E, num_neg, 2] int64) and :6-20 (global_neg_sample -> PyG negative_sampling
// method='sparse': host python random.sample + numpy.isin).  RNG streams of the reference are
// not reproducible (it sets no seed); what is preserved are the distributional contracts:
//   local : src = pos[e,0] repeated num_neg times, dst ~ U[0,N), no filtering.
//   global: DISTINCT cells (r,c) ~ U over {r != c, (r,c) not an existing edge}, in random order.
#include "common.cuh"

namespace plnlp {

__device__ __forceinline__ uint32_t bounded(uint32_t r, uint32_t n) {  // multiply-high map to [0,n)
    return __umulhi(r, n);
}

__global__ void __launch_bounds__(256) local_neg_kernel(const int64_t* __restrict__ pos, int64_t total,
                                                        uint32_t num_nodes, int num_neg, uint64_t seed,
                                                        int64_t* __restrict__ out) {
    // thread t produces outputs 4t .. 4t+3 from one Philox block
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t i0 = t * 4;
    if (i0 >= total) return;
    const uint4 r = philox4x32_10(seed, static_cast<uint64_t>(t), 0x10ca1u);
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int64_t i = i0 + e;
        if (i >= total) break;
        out[2 * i] = __ldg(pos + 2 * (i / num_neg));
        out[2 * i + 1] = static_cast<int64_t>(bounded(rr[e], num_nodes));
    }
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

__device__ __forceinline__ bool sorted_contains(const int64_t* __restrict__ a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int64_t v = __ldg(a + mid);
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(a + lo) == key;
}

constexpr unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFULL;

__global__ void __launch_bounds__(256) global_cand_kernel(const int64_t* __restrict__ edge_ids, int64_t n_edges,
                                                          uint32_t num_nodes, int64_t n_cand, uint64_t seed,
                                                          int64_t* __restrict__ cand_ids,
                                                          unsigned long long* __restrict__ keys,
                                                          unsigned* __restrict__ first, uint64_t mask) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_cand) return;
    const uint4 r = philox4x32_10(seed, static_cast<uint64_t>(i), 0x610ba1u);
    const uint32_t row = bounded(r.x, num_nodes), col = bounded(r.y, num_nodes);
    const int64_t id = static_cast<int64_t>(row) * num_nodes + col;
    const bool valid = (row != col) && !sorted_contains(edge_ids, n_edges, id);
    cand_ids[i] = valid ? id : -1;
    if (!valid) return;
    uint64_t slot = mix64(static_cast<uint64_t>(id)) & mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(keys + slot, kEmpty, static_cast<unsigned long long>(id));
        if (prev == kEmpty || prev == static_cast<unsigned long long>(id)) {
            atomicMin(first + slot, static_cast<unsigned>(i));
            return;
        }
        slot = (slot + 1) & mask;
    }
}

__global__ void __launch_bounds__(256) global_keep_kernel(const int64_t* __restrict__ cand_ids, int64_t n_cand,
                                                          const unsigned long long* __restrict__ keys,
                                                          const unsigned* __restrict__ first, uint64_t mask,
                                                          uint8_t* __restrict__ keep) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_cand) return;
    const int64_t id = cand_ids[i];
    if (id < 0) { keep[i] = 0; return; }
    uint64_t slot = mix64(static_cast<uint64_t>(id)) & mask;
    while (keys[slot] != static_cast<unsigned long long>(id)) slot = (slot + 1) & mask;
    keep[i] = first[slot] == static_cast<unsigned>(i) ? 1 : 0;
}

}  // namespace plnlp

using namespace plnlp;

extern "C" int plnlp_local_neg_sample(const int64_t* pos_edges, int64_t E, int64_t num_nodes, int num_neg,
                                      uint64_t seed, int64_t* out, void* stream) {
    PLNLP_REQUIRE(E >= 0 && num_neg >= 1 && num_nodes > 0 && num_nodes < (1LL << 32), PLNLP_E_SIZE);
    if (E == 0) return 0;
    PLNLP_REQUIRE(pos_edges && out, PLNLP_E_NULL);
    const int64_t total = E * num_neg;
    const unsigned grid = static_cast<unsigned>(ceil_div(ceil_div(total, 4), 256));
    local_neg_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pos_edges, total, static_cast<uint32_t>(num_nodes), num_neg, seed, out);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int plnlp_global_neg_candidates(const int64_t* edge_ids, int64_t n_edges, int64_t num_nodes,
                                           int64_t n_cand, uint64_t seed, int64_t* cand_ids,
                                           unsigned long long* table_keys, int* table_first, int64_t table_size,
                                           void* stream) {
    PLNLP_REQUIRE(n_edges >= 0 && n_cand >= 0 && num_nodes > 0 && num_nodes < (1LL << 31), PLNLP_E_SIZE);
    PLNLP_REQUIRE(n_cand < (1LL << 32) - 1, PLNLP_E_SIZE);
    if (n_cand == 0) return 0;
    PLNLP_REQUIRE(cand_ids && table_keys && table_first && (n_edges == 0 || edge_ids), PLNLP_E_NULL);
    PLNLP_REQUIRE(table_size >= 2 * n_cand && (table_size & (table_size - 1)) == 0, PLNLP_E_SIZE);
    const unsigned grid = static_cast<unsigned>(ceil_div(n_cand, 256));
    global_cand_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        edge_ids, n_edges, static_cast<uint32_t>(num_nodes), n_cand, seed, cand_ids, table_keys,
        reinterpret_cast<unsigned*>(table_first), static_cast<uint64_t>(table_size - 1));
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int plnlp_global_neg_keep(const int64_t* cand_ids, int64_t n_cand, const unsigned long long* table_keys,
                                     const int* table_first, int64_t table_size, uint8_t* keep, void* stream) {
    PLNLP_REQUIRE(n_cand >= 0, PLNLP_E_SIZE);
    if (n_cand == 0) return 0;
    PLNLP_REQUIRE(cand_ids && table_keys && table_first && keep, PLNLP_E_NULL);
    PLNLP_REQUIRE(table_size >= 2 * n_cand && (table_size & (table_size - 1)) == 0, PLNLP_E_SIZE);
    const unsigned grid = static_cast<unsigned>(ceil_div(n_cand, 256));
    global_keep_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        cand_ids, n_cand, table_keys, reinterpret_cast<const unsigned*>(table_first),
        static_cast<uint64_t>(table_size - 1), keep);
    PLNLP_LAUNCH_CHECK();
    return 0;
}
