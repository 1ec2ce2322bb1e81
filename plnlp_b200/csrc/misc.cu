// Elementwise / reduction helpers around the layers, plus ABI bookkeeping.
#include "common.cuh"

namespace plnlp {

long long g_launch_count = 0;

// dx = dy * (y > 0 ? scale : 0)   -- backward of relu (+ inverted dropout) given the forward output
// (/root/reference/plnlp/layer.py:21-22, 25-26, 84-85 via autograd)
template <int VEC>
__global__ void __launch_bounds__(256) relu_drop_bwd_kernel(const float* __restrict__ y, int64_t ldy,
                                                            const float* __restrict__ dy, int64_t lddy, float scale,
                                                            int64_t rows, int64_t cols, float* __restrict__ dx,
                                                            int64_t lddx) {
    const int64_t cv = cols / VEC;
    const int64_t total = rows * cv;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / cv, c = (i % cv) * VEC;
        float a[VEC], g[VEC];
        load_vec<VEC>(a, y + r * ldy + c);
        load_vec<VEC>(g, dy + r * lddy + c);
#pragma unroll
        for (int e = 0; e < VEC; ++e) g[e] = a[e] > 0.0f ? g[e] * scale : 0.0f;
        store_vec<VEC>(dx + r * lddx + c, g);
    }
}

// column sums: stage 1, each block sums CS_RB consecutive rows for 256 columns (coalesced along
// the row, sequential down the rows); stage 2 adds the block partials in block order.
constexpr int CS_RB = 256;

__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows,
                                                             int64_t cols, float* __restrict__ ws) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * CS_RB, r1 = min(rows, r0 + CS_RB);
    if (c >= cols) return;
    // 8 independent row loads in flight per thread (the kernel is a pure HBM stream: with 4 it reached 39 % of
    // the peak); fixed association of the 8 partial sums -> deterministic
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.0f;
    int64_t r = r0;
    for (; r + 7 < r1; r += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(x + (r + i) * ldx + c);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += v[i];
    }
    for (; r < r1; ++r) a[0] += __ldg(x + r * ldx + c);
    ws[static_cast<int64_t>(blockIdx.y) * cols + c] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}

__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ ws, int64_t nblk, int64_t cols,
                                                           float scale, float* __restrict__ out) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (c >= cols) return;
    double acc = 0.0;
    for (int64_t b = 0; b < nblk; ++b) acc += static_cast<double>(ws[b * cols + c]);
    out[c] = static_cast<float>(acc) * scale;
}

// index[r] = any(x[r, :] != 0) ? r : -1: one warp per row, a row-sparsity x_index for the SpMM
template <int VEC>
__global__ void __launch_bounds__(256) row_nonzero_index_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows,
                                                               int F, int32_t* __restrict__ index) {
    const int lane = threadIdx.x & 31;
    const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* xr = x + r * ldx;
    bool nz = false;
    for (int f = lane * VEC; f < F; f += 32 * VEC) {
        float a[VEC];
        load_vec<VEC>(a, xr + f);
#pragma unroll
        for (int e = 0; e < VEC; ++e) nz |= !(a[e] == 0.0f);      // NaN counts as non-zero
    }
    const unsigned any = __ballot_sync(0xffffffffu, nz);
    if (lane == 0) index[r] = any ? static_cast<int32_t>(r) : -1;
}

}  // namespace plnlp


using namespace plnlp;

extern "C" int plnlp_abi_version(void) { return PLNLP_ABI_VERSION; }

extern "C" int64_t plnlp_launch_count(void) { return g_launch_count; }

extern "C" int plnlp_check_device(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PLNLP_E_DEVICE;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return PLNLP_E_DEVICE;
    return major == 10 ? 0 : PLNLP_E_DEVICE;
}

extern "C" int plnlp_relu_drop_bwd_f32(const float* y, int64_t ldy, const float* dy, int64_t lddy, float scale,
                                       int64_t rows, int64_t cols, float* dx, int64_t lddx, void* stream) {
    PLNLP_REQUIRE(rows >= 0 && cols >= 0, PLNLP_E_SIZE);
    if (rows == 0 || cols == 0) return 0;
    PLNLP_REQUIRE(y && dy && dx, PLNLP_E_NULL);
    PLNLP_REQUIRE(ldy >= cols && lddy >= cols && lddx >= cols, PLNLP_E_SIZE);
    const int vec = pick_vec(cols, {ldy, lddy, lddx}, {y, dy, dx});
    const int64_t total = rows * (cols / vec);
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(ceil_div(total, 256), kNumSM * 16));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec == 4)      relu_drop_bwd_kernel<4><<<grid, 256, 0, st>>>(y, ldy, dy, lddy, scale, rows, cols, dx, lddx);
    else if (vec == 2) relu_drop_bwd_kernel<2><<<grid, 256, 0, st>>>(y, ldy, dy, lddy, scale, rows, cols, dx, lddx);
    else               relu_drop_bwd_kernel<1><<<grid, 256, 0, st>>>(y, ldy, dy, lddy, scale, rows, cols, dx, lddx);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t plnlp_colsum_workspace_bytes(int64_t rows, int64_t cols) {
    if (rows < 0 || cols < 0) return 0;
    return std::max<int64_t>(ceil_div(rows, CS_RB), 1) * cols * 4;
}

extern "C" int plnlp_colsum_f32(const float* x, int64_t ldx, int64_t rows, int64_t cols, float scale, float* out,
                                void* workspace, int64_t workspace_bytes, void* stream) {
    PLNLP_REQUIRE(rows >= 0 && cols >= 0, PLNLP_E_SIZE);
    if (cols == 0) return 0;
    PLNLP_REQUIRE(out && workspace && (rows == 0 || x), PLNLP_E_NULL);
    PLNLP_REQUIRE(ldx >= cols, PLNLP_E_SIZE);
    PLNLP_REQUIRE(workspace_bytes >= plnlp_colsum_workspace_bytes(rows, cols), PLNLP_E_WORKSPACE);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t nblk = ceil_div(rows, CS_RB);
    float* ws = static_cast<float*>(workspace);
    const unsigned gx = static_cast<unsigned>(ceil_div(cols, 256));
    if (nblk > 0) {
        colsum_partial_kernel<<<dim3(gx, static_cast<unsigned>(nblk)), 256, 0, st>>>(x, ldx, rows, cols, ws);
        PLNLP_LAUNCH_CHECK();
    }
    colsum_final_kernel<<<gx, 256, 0, st>>>(ws, nblk, cols, scale, out);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int plnlp_row_nonzero_index_f32(const float* x, int64_t ldx, int64_t rows, int64_t F, int32_t* index,
                                           void* stream) {
    using namespace plnlp;
    PLNLP_REQUIRE(rows >= 0 && rows < (1ll << 31) && F > 0 && ldx >= F, PLNLP_E_SIZE);
    if (rows == 0) return 0;
    PLNLP_REQUIRE(x && index, PLNLP_E_NULL);
    const int vec = pick_vec(F, {ldx}, {x});
    const unsigned grid = static_cast<unsigned>(ceil_div(rows, 8));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec == 4) row_nonzero_index_kernel<4><<<grid, 256, 0, st>>>(x, ldx, rows, static_cast<int>(F), index);
    else if (vec == 2) row_nonzero_index_kernel<2><<<grid, 256, 0, st>>>(x, ldx, rows, static_cast<int>(F), index);
    else row_nonzero_index_kernel<1><<<grid, 256, 0, st>>>(x, ldx, rows, static_cast<int>(F), index);
    PLNLP_LAUNCH_CHECK();
    return 0;
}
