// Pairwise ranking losses and d loss / d score in one pass.
//
// Replaces /root/reference/plnlp/loss.py:5-8 (auc_loss), 11-14 (hinge_auc_loss), 31-35
// (weighted_hinge_auc_loss) and the ~12 elementwise launches of their autograd mirror.
//   t_ij = m_i - (pos_i - neg_ij),   m_i = 1 (AUC, HingeAUC) or w_i (WeightedHingeAUC)
//   AUC:              loss = sum t^2              d/dneg_ij = 2 t,           d/dpos_i = -sum_j
//   HingeAUC:         loss = sum max(t,0)^2       d/dneg_ij = 2 max(t,0)
//   WeightedHingeAUC: loss = sum w_i max(t,0)^2   d/dneg_ij = 2 w_i max(t,0)
// The loss is a SUM (not a mean), as in the reference.  Block partials are accumulated in
// fp64 and reduced in block order: deterministic.
#include "common.cuh"

namespace plnlp {

constexpr int LOSS_TB = 256;

__global__ void __launch_bounds__(LOSS_TB) pair_loss_kernel(int kind, const float* __restrict__ pos,
                                                            const float* __restrict__ neg,
                                                            const float* __restrict__ weight, int64_t B, int k,
                                                            float* __restrict__ dpos, float* __restrict__ dneg,
                                                            double* __restrict__ ws) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * LOSS_TB + threadIdx.x;
    double l = 0.0;
    if (i < B) {
        const float p = __ldg(pos + i);
        const float w = (kind == PLNLP_LOSS_WEIGHTED_HINGE_AUC) ? __ldg(weight + i) : 1.0f;
        float gp = 0.0f;
        for (int j = 0; j < k; ++j) {
            const float n = __ldg(neg + i * k + j);
            float t = w - (p - n);
            if (kind != PLNLP_LOSS_AUC) t = fmaxf(t, 0.0f);
            const float tt = t * t;
            l += static_cast<double>(kind == PLNLP_LOSS_WEIGHTED_HINGE_AUC ? w * tt : tt);
            float g = 2.0f * t;
            if (kind == PLNLP_LOSS_WEIGHTED_HINGE_AUC) g *= w;
            dneg[i * k + j] = g;
            gp -= g;
        }
        dpos[i] = gp;
    }
    __shared__ double sm[LOSS_TB / 32];
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w8 = 0; w8 < LOSS_TB / 32; ++w8) s += sm[w8];
        ws[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) pair_loss_final_kernel(const double* __restrict__ ws, int64_t nblk,
                                                              float* __restrict__ loss) {
    __shared__ double sm[256];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < nblk; i += 256) acc += ws[i];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (static_cast<int>(threadIdx.x) < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = static_cast<float>(sm[0]);
}

}  // namespace plnlp

using namespace plnlp;

extern "C" int64_t plnlp_pair_loss_workspace_bytes(int64_t B) {
    if (B < 0) return 0;
    return (ceil_div(B, LOSS_TB) + 1) * 8;
}

extern "C" int plnlp_pair_loss_f32(int kind, const float* pos, const float* neg, const float* weight, int64_t B,
                                   int num_neg, float* loss, float* dpos, float* dneg, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
    PLNLP_REQUIRE(kind >= 0 && kind <= 2, PLNLP_E_UNSUPPORTED);
    PLNLP_REQUIRE(B >= 0 && num_neg >= 1, PLNLP_E_SIZE);
    PLNLP_REQUIRE(loss && workspace, PLNLP_E_NULL);
    if (B > 0) PLNLP_REQUIRE(pos && neg && dpos && dneg, PLNLP_E_NULL);
    if (kind == PLNLP_LOSS_WEIGHTED_HINGE_AUC && B > 0) PLNLP_REQUIRE(weight, PLNLP_E_NULL);
    PLNLP_REQUIRE(workspace_bytes >= plnlp_pair_loss_workspace_bytes(B), PLNLP_E_WORKSPACE);
    PLNLP_REQUIRE(aligned(workspace, 8), PLNLP_E_ALIGN);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t nblk = ceil_div(B, LOSS_TB);
    double* ws = static_cast<double*>(workspace);
    if (nblk > 0) {
        pair_loss_kernel<<<static_cast<unsigned>(nblk), LOSS_TB, 0, st>>>(kind, pos, neg, weight, B, num_neg, dpos,
                                                                          dneg, ws);
        PLNLP_LAUNCH_CHECK();
    }
    pair_loss_final_kernel<<<1, 256, 0, st>>>(ws, nblk, loss);
    PLNLP_LAUNCH_CHECK();
    return 0;
}
