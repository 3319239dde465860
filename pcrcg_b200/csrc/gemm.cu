// Dense contractions of the hot path: C[M,N] = A[M,K] * B (+ per-row scale).
//   * KPConv weight contraction   [Nq, K*Cin] x [K*Cin, Cout]      models/blocks.py:361-366
//   * UnaryBlock / shortcut Linear [N, Cin]   x [Cout, Cin]^T       models/blocks.py:490,497
// gemm_dev dispatches to the tcgen05 tensor-core kernel (gemm_tc.cu) when the shape qualifies and
// to the fp32 CUDA-core kernel below otherwise (ragged K/N such as Cin = 1 or 129).
#include "common.cuh"

namespace pcrcg {

constexpr int GB_M = 64, GB_N = 64, GB_K = 16;

// B element (k,n): b_is_nk ? B[n*ldb + k] : B[k*ldb + n]
template <bool B_NK>
__global__ void __launch_bounds__(256) k_sgemm(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                               float* __restrict__ C, int ldc, int M, int N, int K,
                                               const float* __restrict__ row_scale)
{
    __shared__ float As[GB_K][GB_M + 4];
    __shared__ float Bs[GB_K][GB_N + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
    const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, 4x4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += GB_K) {
        // A tile: 64 rows x 16 k ; thread loads 4 elements, k fastest (coalesced along lda rows)
#pragma unroll
        for (int t = 0; t < 4; t++) {
            int e = tid + t * 256;
            int r = e >> 4, kk = e & 15;
            int gm = m0 + r, gk = k0 + kk;
            As[kk][r] = (gm < M && gk < K) ? A[(size_t)gm * lda + gk] : 0.f;
        }
        if (B_NK) {
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int e = tid + t * 256;
                int c = e >> 4, kk = e & 15;
                int gn = n0 + c, gk = k0 + kk;
                Bs[kk][c] = (gn < N && gk < K) ? B[(size_t)gn * ldb + gk] : 0.f;
            }
        } else {
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int e = tid + t * 256;
                int kk = e >> 6, c = e & 63;
                int gn = n0 + c, gk = k0 + kk;
                Bs[kk][c] = (gn < N && gk < K) ? B[(size_t)gk * ldb + gn] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GB_K; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
        float s = row_scale ? row_scale[gm] : 1.0f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int gn = n0 + tx * 4 + j;
            if (gn < N) C[(size_t)gm * ldc + gn] = acc[i][j] * s;
        }
    }
}

int gemm_tc_dev(const float* A, int lda, const float* B, int ldb, int b_is_nk, float* C, int ldc, int M, int N, int K,
                const float* row_scale, cudaStream_t st, bool* handled);   // gemm_tc.cu

static int g_force_simt = 0;
void gemm_set_force_simt(int v) { g_force_simt = v; }
int gemm_force_simt_get() { return g_force_simt; }
// profile class of gemm_dev's launches: the KPConv weight contraction (PC_GEMM, the callers inside kpconv.cu) or a unary
// Linear / any other dense product entered through the C ABI (PC_LINEAR, set by capi.cu around the call)
static int g_gemm_prof_class = PC_GEMM;
void gemm_set_prof_class(int c) { g_gemm_prof_class = c; }

int gemm_dev(const float* A, int lda, const float* B, int ldb, int b_is_nk, float* C, int ldc, int M, int N, int K,
             const float* row_scale, cudaStream_t st)
{
    if (M <= 0 || N <= 0) return PCRCG_OK;
    PCRCG_REQUIRE(K >= 1, "gemm: K must be >= 1");
    ProfScope prof(g_gemm_prof_class, st, 1);
    if (!g_force_simt) {
        bool handled = false;
        PCRCG_TRY(gemm_tc_dev(A, lda, B, ldb, b_is_nk, C, ldc, M, N, K, row_scale, st, &handled));
        if (handled) return PCRCG_OK;
    }
    // M tiles ride on grid.y (limit 65535): more than 65535 x 64 rows (a stacked batch beyond ~4.19 M points reaches this
    // kernel through the Cin = 1 first layer) go in several launches
    const int rows_per_launch = 65535 * GB_M;
    for (int m0 = 0; m0 < M; m0 += rows_per_launch) {
        const int mc = M - m0 < rows_per_launch ? M - m0 : rows_per_launch;
        const float* a = A + (size_t)m0 * lda;
        float* c = C + (size_t)m0 * ldc;
        const float* rs = row_scale ? row_scale + m0 : nullptr;
        dim3 grid((unsigned)cdiv64(N, GB_N), (unsigned)cdiv64(mc, GB_M));
        if (m0 > 0) count_launches(1);
        if (b_is_nk) k_sgemm<true><<<grid, 256, 0, st>>>(a, lda, B, ldb, c, ldc, mc, N, K, rs);
        else k_sgemm<false><<<grid, 256, 0, st>>>(a, lda, B, ldb, c, ldc, mc, N, K, rs);
    }
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

}  // namespace pcrcg
