// Ranking metrics on the GPU.
//
// Replaces ogb.linkproppred.Evaluator._eval_hits / _eval_mrr running on CPU tensors
// (/root/reference/plnlp/utils.py:49-56 and 67-76):
//   Hits@K: kth = K-th largest negative score; hits = #{pos > kth} / #pos  (strict >)
//   MRR   : rank_r = 1 + #{neg_r > pos_r} (optimistic) ; ge counts are returned too so callers can
//           detect ties, where ogb 1.3.2's argsort-based rank is implementation defined.
// All integer work: bit-exact and deterministic.
#include "common.cuh"

namespace plnlp {

struct SelectState {
    uint32_t prefix;
    uint32_t pad;
    unsigned long long k_rem;
};

// one MSB-first 8-bit radix-select pass: histogram of the digit at `shift` over keys whose higher
// digits equal the prefix chosen so far
__global__ void __launch_bounds__(256) select_hist_kernel(const float* __restrict__ x, int64_t n, int pass,
                                                          const SelectState* __restrict__ st,
                                                          unsigned* __restrict__ hist) {
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int shift = 24 - 8 * pass;
    const uint32_t prefix = st->prefix;
    const uint32_t himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint32_t k = float_key(__ldg(x + i));
        if ((k & himask) == (prefix & himask)) atomicAdd(&sh[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(hist + pass * 256 + threadIdx.x, sh[threadIdx.x]);
}

__global__ void select_pick_kernel(int pass, SelectState* __restrict__ st, const unsigned* __restrict__ hist,
                                   float* __restrict__ kth) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int shift = 24 - 8 * pass;
    unsigned long long k = st->k_rem, cum = 0;
    int bin = 255;
    for (; bin > 0; --bin) {
        const unsigned h = hist[pass * 256 + bin];
        if (cum + h >= k) break;
        cum += h;
    }
    st->k_rem = k - cum;
    st->prefix |= static_cast<uint32_t>(bin) << shift;
    if (pass == 3) kth[0] = key_float(st->prefix);
}

__global__ void select_init_kernel(SelectState* st, unsigned* hist, unsigned long long K) {
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x == 0) { st->prefix = 0; st->pad = 0; st->k_rem = K; }
}

__global__ void __launch_bounds__(256) count_greater_kernel(const float* __restrict__ pos, int64_t n,
                                                            const float* __restrict__ thresh,
                                                            unsigned long long* __restrict__ count) {
    const float t = __ldg(thresh);
    unsigned c = 0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        c += (__ldg(pos + i) > t) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, static_cast<unsigned long long>(c));
}

__global__ void zero_u64_kernel(unsigned long long* p) { p[0] = 0; }

__global__ void __launch_bounds__(256) mrr_counts_kernel(const float* __restrict__ pos, const float* __restrict__ neg,
                                                         int64_t ldn, int64_t S, int64_t K, int32_t* __restrict__ gt,
                                                         int32_t* __restrict__ ge) {
    const int lane = threadIdx.x & 31;
    const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= S) return;
    const float p = __ldg(pos + r);
    const float* nr = neg + r * ldn;
    int cgt = 0, cge = 0;
    for (int64_t j = lane; j < K; j += 32) {
        const float v = __ldg(nr + j);
        cgt += v > p;
        cge += v >= p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cgt += __shfl_xor_sync(0xffffffffu, cgt, o);
        cge += __shfl_xor_sync(0xffffffffu, cge, o);
    }
    if (lane == 0) { gt[r] = cgt; ge[r] = cge; }
}

}  // namespace plnlp

using namespace plnlp;

extern "C" int plnlp_kth_largest_f32(const float* neg, int64_t n, int64_t K, float* kth, void* workspace,
                                     int64_t workspace_bytes, void* stream) {
    PLNLP_REQUIRE(n >= 1 && K >= 1 && K <= n, PLNLP_E_SIZE);
    PLNLP_REQUIRE(neg && kth && workspace, PLNLP_E_NULL);
    PLNLP_REQUIRE(workspace_bytes >= 8192, PLNLP_E_WORKSPACE);
    PLNLP_REQUIRE(aligned(workspace, 16), PLNLP_E_ALIGN);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned* hist = static_cast<unsigned*>(workspace);
    SelectState* state = reinterpret_cast<SelectState*>(static_cast<char*>(workspace) + 4096);
    select_init_kernel<<<1, 256, 0, st>>>(state, hist, static_cast<unsigned long long>(K));
    PLNLP_LAUNCH_CHECK();
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(ceil_div(n, 256 * 8), kNumSM * 8));
    for (int pass = 0; pass < 4; ++pass) {
        select_hist_kernel<<<grid ? grid : 1, 256, 0, st>>>(neg, n, pass, state, hist);
        PLNLP_LAUNCH_CHECK();
        select_pick_kernel<<<1, 32, 0, st>>>(pass, state, hist, kth);
        PLNLP_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int plnlp_count_greater_f32(const float* pos, int64_t n, const float* thresh, unsigned long long* count,
                                       void* stream) {
    PLNLP_REQUIRE(n >= 0, PLNLP_E_SIZE);
    PLNLP_REQUIRE(thresh && count && (n == 0 || pos), PLNLP_E_NULL);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    zero_u64_kernel<<<1, 1, 0, st>>>(count);
    PLNLP_LAUNCH_CHECK();
    if (n == 0) return 0;
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(ceil_div(n, 256 * 4), kNumSM * 8));
    count_greater_kernel<<<grid ? grid : 1, 256, 0, st>>>(pos, n, thresh, count);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int plnlp_mrr_counts_f32(const float* pos, const float* neg, int64_t ldn, int64_t S, int64_t K,
                                    int32_t* gt, int32_t* ge, void* stream) {
    PLNLP_REQUIRE(S >= 0 && K >= 0 && ldn >= K, PLNLP_E_SIZE);
    if (S == 0) return 0;
    PLNLP_REQUIRE(pos && gt && ge && (K == 0 || neg), PLNLP_E_NULL);
    const unsigned grid = static_cast<unsigned>(ceil_div(S, 8));
    mrr_counts_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(pos, neg, ldn, S, K, gt, ge);
    PLNLP_LAUNCH_CHECK();
    return 0;
}
