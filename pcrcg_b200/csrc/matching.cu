// Descriptor matching front-end ("next" row 4 of the scope table) -- sm_100a.
//
// Replaces the dense score matrix of lib/benchmark_utils.py:187-224,246-262 (`scores = src_feat @ tgt_feat^T`, then
// argmax along rows / columns, mutual_selection :270-295) by a fused kernel that never materialises the [N, M] matrix:
// one thread per row of A keeps that row in registers, B is streamed through shared memory in tiles, and the running
// (best score, first index reaching it) is kept per thread -- np.argmax's first-maximum rule.
#include "common.cuh"

namespace pcrcg {

constexpr int BM_TILE = 64;      // rows of B per shared-memory tile

template <int D>
__global__ void __launch_bounds__(128) k_best_match(const float* __restrict__ a, int n, const float* __restrict__ b, int m,
                                                    int32_t* __restrict__ best_idx, float* __restrict__ best_val)
{
    __shared__ float s_b[BM_TILE][D];
    const int i = blockIdx.x * 128 + threadIdx.x;
    float row[D];
#pragma unroll
    for (int d = 0; d < D; d++) row[d] = i < n ? a[(size_t)i * D + d] : 0.f;
    float bv = -INFINITY;
    int bi = 0;
    for (int j0 = 0; j0 < m; j0 += BM_TILE) {
        const int cnt = min(BM_TILE, m - j0);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * D; e += 128) s_b[e / D][e % D] = b[(size_t)j0 * D + e];
        __syncthreads();
        for (int j = 0; j < cnt; j++) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < D; d++) s = fmaf(row[d], s_b[j][d], s);
            if (s > bv) { bv = s; bi = j0 + j; }         // strict: the first maximum wins
        }
    }
    if (i < n) {
        best_idx[i] = bi;
        if (best_val != nullptr) best_val[i] = bv;
    }
}

// mutual[i] = (col_best[row_best[i]] == i)      lib/benchmark_utils.py:270-295
__global__ void __launch_bounds__(256) k_mutual(const int32_t* __restrict__ row_best, const int32_t* __restrict__ col_best, int n,
                                                uint8_t* __restrict__ mutual)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mutual[i] = col_best[row_best[i]] == i ? 1 : 0;
}

int best_match_dev(const float* a, int64_t n, const float* b, int64_t m, int32_t D, int32_t* best_idx, float* best_val, cudaStream_t st)
{
    PCRCG_REQUIRE(n >= 0 && m >= 1 && n < (1ll << 31) && m < (1ll << 31), "best_match: bad sizes");
    PCRCG_REQUIRE(D == 16 || D == 32 || D == 64, "best_match: descriptor length must be 16, 32 or 64");
    if (n == 0) return PCRCG_OK;
    ProfScope prof(PC_GEMM, st, 1);
    const unsigned g = (unsigned)cdiv64(n, 128);
    if (D == 16) k_best_match<16><<<g, 128, 0, st>>>(a, (int)n, b, (int)m, best_idx, best_val);
    else if (D == 32) k_best_match<32><<<g, 128, 0, st>>>(a, (int)n, b, (int)m, best_idx, best_val);
    else k_best_match<64><<<g, 128, 0, st>>>(a, (int)n, b, (int)m, best_idx, best_val);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int mutual_dev(const int32_t* row_best, const int32_t* col_best, int64_t n, uint8_t* mutual, cudaStream_t st)
{
    if (n <= 0) return PCRCG_OK;
    ProfScope prof(PC_POOL, st, 1);
    k_mutual<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(row_best, col_best, (int)n, mutual);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

}  // namespace pcrcg
