// Uniform random walks + (start, visited) pair expansion on the GPU.
//
// Replaces torch_cluster.random_walk and the python pair / weight assembly that produces the training
// pairs of the random-walk augmentation (/root/reference/main.py:228-233, 241-253):
//   walk[n, 0] = start[n];  walk[n, l+1] = col[rowptr[cur] + floor(u * deg(cur))]  (stay put if deg == 0)
//   pairs  = for j in 0..L-1: (walk[:, 0], walk[:, j+1])  (j-major order), weight 1/(j+1),
//            self pairs (src == dst) flagged for removal.
// Uniforms come from Philox4x32-10 (seeded) or, for bit-exact parity tests against the CPU restatement,
// from a caller-supplied [n_walks, L] array -- upstream also pre-draws torch.rand(n_walks, L).
#include "common.cuh"

namespace plnlp {

__global__ void __launch_bounds__(256) random_walk_kernel(const int64_t* __restrict__ rowptr,
                                                          const int64_t* __restrict__ col,
                                                          const int64_t* __restrict__ start, int64_t n_walks, int L,
                                                          const float* __restrict__ rand, uint64_t seed,
                                                          int64_t* __restrict__ walk) {
    const int64_t n = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (n >= n_walks) return;
    int64_t cur = __ldg(start + n);
    int64_t* w = walk + n * (L + 1);
    w[0] = cur;
    uint4 r = make_uint4(0, 0, 0, 0);
    for (int l = 0; l < L; ++l) {
        float u;
        if (rand) {
            u = __ldg(rand + n * L + l);
        } else {
            if ((l & 3) == 0) r = philox4x32_10(seed, static_cast<uint64_t>(n) * ((L + 3) / 4) + (l >> 2), 0x3a1cu);
            const uint32_t word = (l & 3) == 0 ? r.x : (l & 3) == 1 ? r.y : (l & 3) == 2 ? r.z : r.w;
            u = u32_to_unit(word);
        }
        const int64_t b = __ldg(rowptr + cur), e = __ldg(rowptr + cur + 1);
        const int64_t deg = e - b;
        if (deg > 0) {
            int64_t off = static_cast<int64_t>(u * static_cast<float>(deg));
            if (off >= deg) off = deg - 1;      // u*deg can round up to deg in fp32
            cur = __ldg(col + b + off);
        }
        w[l + 1] = cur;
    }
}

// pairs[j*W + n] = (walk[n,0], walk[n,j+1]); weight = 1/(j+1); keep = src != dst
__global__ void __launch_bounds__(256) walk_pairs_kernel(const int64_t* __restrict__ walk, int64_t n_walks, int L,
                                                         int64_t* __restrict__ pairs, float* __restrict__ weight,
                                                         uint8_t* __restrict__ keep) {
    const int64_t total = n_walks * L;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t j = i / n_walks, n = i - j * n_walks;
        const int64_t s = __ldg(walk + n * (L + 1)), d = __ldg(walk + n * (L + 1) + j + 1);
        pairs[2 * i] = s;
        pairs[2 * i + 1] = d;
        weight[i] = 1.0f / static_cast<float>(j + 1);
        keep[i] = s != d;
    }
}

}  // namespace plnlp

using namespace plnlp;

extern "C" int plnlp_random_walk(const int64_t* rowptr, const int64_t* col, const int64_t* start, int64_t n_walks,
                                 int walk_length, const float* rand, uint64_t seed, int64_t* walk, void* stream) {
    PLNLP_REQUIRE(n_walks >= 0 && walk_length >= 0, PLNLP_E_SIZE);
    if (n_walks == 0) return 0;
    PLNLP_REQUIRE(rowptr && start && walk && (col || walk_length == 0), PLNLP_E_NULL);
    const unsigned grid = static_cast<unsigned>(ceil_div(n_walks, 256));
    random_walk_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(rowptr, col, start, n_walks, walk_length,
                                                                           rand, seed, walk);
    PLNLP_LAUNCH_CHECK();
    return 0;
}

extern "C" int plnlp_walk_pairs(const int64_t* walk, int64_t n_walks, int walk_length, int64_t* pairs, float* weight,
                                uint8_t* keep, void* stream) {
    PLNLP_REQUIRE(n_walks >= 0 && walk_length >= 0, PLNLP_E_SIZE);
    if (n_walks == 0 || walk_length == 0) return 0;
    PLNLP_REQUIRE(walk && pairs && weight && keep, PLNLP_E_NULL);
    const int64_t total = n_walks * walk_length;
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(ceil_div(total, 256), kNumSM * 16));
    walk_pairs_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(walk, n_walks, walk_length, pairs, weight,
                                                                          keep);
    PLNLP_LAUNCH_CHECK();
    return 0;
}
