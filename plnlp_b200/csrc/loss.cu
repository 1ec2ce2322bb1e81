// Pairwise ranking losses and d loss / d score in one pass.
//
// Replaces /root/reference/plnlp/loss.py and the ~12 elementwise launches of each loss's autograd
// mirror.  With x_ij = pos_i - neg_ij, m_i the margin and w_i the weight:
//   AUC              (loss.py:5-8)    sum (1 - x)^2
//   HingeAUC         (loss.py:11-14)  sum max(0, 1 - x)^2
//   WeightedAUC      (loss.py:17-21)  sum w_i (1 - x)^2
//   AdaAUC           (loss.py:24-28)  sum (m_i - x)^2
//   WeightedHingeAUC (loss.py:31-35)  sum w_i max(0, w_i - x)^2        (w is weight AND margin)
//   AdaHingeAUC      (loss.py:38-42)  sum max(0, m_i - x)^2
//   LogRank          (loss.py:45-48)  mean -log(sigmoid(x) + 1e-15)
//   CE               (loss.py:51-54)  mean_i -log(sigmoid(pos_i) + 1e-15) + mean_ij -log(1 - sigmoid(neg_ij) + 1e-15)
//   InfoNCE          (loss.py:57-62)  mean_i -log(e^pos_i / (e^pos_i + sum_j e^neg_ij) + 1e-15)
// The AUC family is a SUM, the last three are MEANS, as in the reference.  Block partials are accumulated
// in fp64 and reduced in block order: deterministic.
#include "common.cuh"

namespace plnlp {

constexpr int LOSS_TB = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(LOSS_TB) pair_loss_kernel(int kind, const float* __restrict__ pos,
                                                            const float* __restrict__ neg,
                                                            const float* __restrict__ weight, int64_t B, int k,
                                                            float* __restrict__ dpos, float* __restrict__ dneg,
                                                            double* __restrict__ ws) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * LOSS_TB + threadIdx.x;
    double l = 0.0;
    if (i < B) {
        const float p = __ldg(pos + i);
        float gp = 0.0f;
        if (kind <= PLNLP_LOSS_ADA_HINGE_AUC) {
            // squared-margin family: t = margin - x, optional hinge, optional weight
            const float wi = weight ? __ldg(weight + i) : 1.0f;
            const bool hinge = kind == PLNLP_LOSS_HINGE_AUC || kind == PLNLP_LOSS_WEIGHTED_HINGE_AUC ||
                               kind == PLNLP_LOSS_ADA_HINGE_AUC;
            const bool weighted = kind == PLNLP_LOSS_WEIGHTED_AUC || kind == PLNLP_LOSS_WEIGHTED_HINGE_AUC;
            const bool margin_is_w = kind == PLNLP_LOSS_ADA_AUC || kind == PLNLP_LOSS_WEIGHTED_HINGE_AUC ||
                                     kind == PLNLP_LOSS_ADA_HINGE_AUC;
            const float m = margin_is_w ? wi : 1.0f;
            const float sc = weighted ? wi : 1.0f;
            for (int j = 0; j < k; ++j) {
                const float n = __ldg(neg + i * k + j);
                float t = m - (p - n);
                if (hinge) t = fmaxf(t, 0.0f);
                l += static_cast<double>(sc * (t * t));
                const float g = 2.0f * sc * t;
                dneg[i * k + j] = g;
                gp -= g;
            }
        } else if (kind == PLNLP_LOSS_LOG_RANK) {
            const float inv = 1.0f / (static_cast<float>(B) * static_cast<float>(k));
            for (int j = 0; j < k; ++j) {
                const float n = __ldg(neg + i * k + j);
                const float s = sigmoidf_(p - n);
                l += static_cast<double>(-logf(s + 1e-15f) * inv);
                const float dx = -(s * (1.0f - s)) / (s + 1e-15f) * inv;      // d loss / d x
                gp += dx;
                dneg[i * k + j] = -dx;
            }
        } else if (kind == PLNLP_LOSS_CE) {
            const float invp = 1.0f / static_cast<float>(B), invn = 1.0f / (static_cast<float>(B) * k);
            const float sp = sigmoidf_(p);
            l += static_cast<double>(-logf(sp + 1e-15f) * invp);
            gp = -(sp * (1.0f - sp)) / (sp + 1e-15f) * invp;
            for (int j = 0; j < k; ++j) {
                const float n = __ldg(neg + i * k + j);
                const float sn = sigmoidf_(n);
                l += static_cast<double>(-logf(1.0f - sn + 1e-15f) * invn);
                dneg[i * k + j] = (sn * (1.0f - sn)) / (1.0f - sn + 1e-15f) * invn;
            }
        } else {  // InfoNCE
            const float inv = 1.0f / static_cast<float>(B);
            const float ep = expf(p);
            float se = 0.0f;
            for (int j = 0; j < k; ++j) se += expf(__ldg(neg + i * k + j));
            const float r = ep / (ep + se);
            l += static_cast<double>(-logf(r + 1e-15f) * inv);
            const float dr = -inv / (r + 1e-15f);                           // d loss / d r
            gp = dr * (ep * se) / ((ep + se) * (ep + se));                   // dr/dp = ep*se/(ep+se)^2
            for (int j = 0; j < k; ++j) {
                const float en = expf(__ldg(neg + i * k + j));
                dneg[i * k + j] = dr * (-ep * en) / ((ep + se) * (ep + se));
            }
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
    PLNLP_REQUIRE(kind >= 0 && kind <= PLNLP_LOSS_INFO_NCE, PLNLP_E_UNSUPPORTED);
    PLNLP_REQUIRE(B >= 0 && num_neg >= 1, PLNLP_E_SIZE);
    PLNLP_REQUIRE(loss && workspace, PLNLP_E_NULL);
    if (B > 0) PLNLP_REQUIRE(pos && neg && dpos && dneg, PLNLP_E_NULL);
    const bool needs_w = kind == PLNLP_LOSS_WEIGHTED_HINGE_AUC || kind == PLNLP_LOSS_WEIGHTED_AUC ||
                         kind == PLNLP_LOSS_ADA_AUC || kind == PLNLP_LOSS_ADA_HINGE_AUC;
    if (needs_w && B > 0) PLNLP_REQUIRE(weight, PLNLP_E_NULL);
    PLNLP_REQUIRE(workspace_bytes >= plnlp_pair_loss_workspace_bytes(B), PLNLP_E_WORKSPACE);
    PLNLP_REQUIRE(aligned(workspace, 8), PLNLP_E_ALIGN);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t nblk = ceil_div(B, LOSS_TB);
    double* ws = static_cast<double*>(workspace);
    if (nblk > 0) {
        pair_loss_kernel<<<static_cast<unsigned>(nblk), LOSS_TB, 0, st>>>(kind, pos, neg, needs_w ? weight : nullptr,
                                                                          B, num_neg, dpos, dneg, ws);
        PLNLP_LAUNCH_CHECK();
    }
    pair_loss_final_kernel<<<1, 256, 0, st>>>(ws, nblk, loss);
    PLNLP_LAUNCH_CHECK();
    return 0;
}
