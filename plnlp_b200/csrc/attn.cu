// Per-destination softmax over the stored entries of a CSR row (forward and backward): the attention weights of
// PyG's TransformerConv, the conv of the reference's `Transformer` encoder (/root/reference/plnlp/layer.py:57-63).
//   alpha[e] = exp(s[e] - max_row) / sum_row exp(s - max_row)           for e in [rowptr[r], rowptr[r+1])
//   ds[e]    = alpha[e] * (dalpha[e] - sum_row alpha * dalpha)
// One warp per row; entries are walked 32 at a time, reductions by shuffle in a fixed order (deterministic).
// Latency bound (8 B per entry); the heavy parts of the conv are the edge-dot and SpMM kernels.
#include "common.cuh"

namespace plnlp {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(256) segment_softmax_fwd_kernel(const int64_t* __restrict__ rowptr, int64_t n_rows,
                                                                  const float* __restrict__ s, float scale,
                                                                  float* __restrict__ alpha) {
    const int lane = threadIdx.x & 31;
    const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    const int64_t e0 = __ldg(rowptr + r), e1 = __ldg(rowptr + r + 1);
    float m = -INFINITY;
    for (int64_t e = e0 + lane; e < e1; e += 32) m = fmaxf(m, __ldg(s + e) * scale);
    m = warp_max(m);
    float z = 0.0f;
    for (int64_t e = e0 + lane; e < e1; e += 32) z += expf(__ldg(s + e) * scale - m);
    z = warp_sum(z);
    const float inv = 1.0f / z;
    for (int64_t e = e0 + lane; e < e1; e += 32) alpha[e] = expf(__ldg(s + e) * scale - m) * inv;
}

__global__ void __launch_bounds__(256) segment_softmax_bwd_kernel(const int64_t* __restrict__ rowptr, int64_t n_rows,
                                                                  const float* __restrict__ alpha,
                                                                  const float* __restrict__ dalpha, float scale,
                                                                  float* __restrict__ ds) {
    const int lane = threadIdx.x & 31;
    const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    const int64_t e0 = __ldg(rowptr + r), e1 = __ldg(rowptr + r + 1);
    float dot = 0.0f;
    for (int64_t e = e0 + lane; e < e1; e += 32) dot = fmaf(__ldg(alpha + e), __ldg(dalpha + e), dot);
    dot = warp_sum(dot);
    for (int64_t e = e0 + lane; e < e1; e += 32) ds[e] = __ldg(alpha + e) * (__ldg(dalpha + e) - dot) * scale;
}

}  // namespace plnlp

extern "C" int plnlp_segment_softmax_fwd_f32(const int64_t* rowptr, int64_t n_rows, const float* s, float scale,
                                             float* alpha, void* stream) {
    using namespace plnlp;
    PLNLP_REQUIRE(n_rows >= 0, PLNLP_E_SIZE);
    if (n_rows == 0) return 0;
    PLNLP_REQUIRE(rowptr && s && alpha, PLNLP_E_NULL);
    segment_softmax_fwd_kernel<<<static_cast<unsigned>(ceil_div(n_rows, 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        rowptr, n_rows, s, scale, alpha);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int plnlp_segment_softmax_bwd_f32(const int64_t* rowptr, int64_t n_rows, const float* alpha,
                                             const float* dalpha, float scale, float* ds, void* stream) {
    using namespace plnlp;
    PLNLP_REQUIRE(n_rows >= 0, PLNLP_E_SIZE);
    if (n_rows == 0) return 0;
    PLNLP_REQUIRE(rowptr && alpha && dalpha && ds, PLNLP_E_NULL);
    segment_softmax_bwd_kernel<<<static_cast<unsigned>(ceil_div(n_rows, 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        rowptr, n_rows, alpha, dalpha, scale, ds);
    PLNLP_LAUNCH_CHECK();
    return 0;
}
