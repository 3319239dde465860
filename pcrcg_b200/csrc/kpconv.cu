// KPConv forward (rigid kernel, linear influence, sum aggregation) -- sm_100a.
//
// Replaces models/blocks.py:229-374 (KPConv.forward, non-deformable branch).  Math:
//   w[n,k,h] = max(0, 1 - |s[idx[n,h]] - q[n] - kp[k]| / KP_extent)          (shadow neighbour -> 0)
//   wf[n,k,:] = sum_h w[n,k,h] * x[idx[n,h],:]
//   out[n,:]  = (sum_k wf[n,k,:] @ W[k]) / max(1, #{h : sum_c x[idx[n,h],c] > 0})
//
// Stage 1 (this file, k_kpconv_aggregate): one warp per query point, lanes over channels.  For each
// neighbour the 15 influence weights are computed once by lanes 0..14, published through shared
// memory and applied to the coalesced feature row; the [K*Cin] aggregate stays in registers and is
// written once.  Stage 2 is the [Nq, K*Cin] x [K*Cin, Cout] contraction (gemm.cu) with the
// 1/count row scale in its epilogue.
#include "common.cuh"

#include <cuda_bf16.h>

namespace pcrcg {

constexpr int KP_MAX = 16;          // kernel points padded to 16
constexpr int AGG_WARPS = 8;

// flag[s] = (sum_c x[s,c] > 0)      models/blocks.py:369-370 (per support row, shared by all queries)
__global__ void __launch_bounds__(256) k_row_positive(const float* __restrict__ x, int n, int c, int ldx, uint8_t* __restrict__ flag)
{
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    float s = 0.f;
    for (int k = lane; k < c; k += 32) s += x[(size_t)row * ldx + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) flag[row] = s > 0.f ? 1 : 0;
}

template <typename IdxT, int CJ, bool SPLIT>
__global__ void __launch_bounds__(AGG_WARPS * 32) k_kpconv_aggregate(
    const float* __restrict__ q_pts, int nq, const float* __restrict__ s_pts, int ns, const IdxT* __restrict__ idx, int H,
    int idx_stride, const float* __restrict__ x, int cin, int ldx, const uint8_t* __restrict__ rowflag,
    const float* __restrict__ kpts, int K, float inv_extent, float* __restrict__ wf, __nv_bfloat16* __restrict__ wf_hi,
    __nv_bfloat16* __restrict__ wf_lo, int ldk, float* __restrict__ inv_cnt)
{
    __shared__ __align__(16) float s_w[AGG_WARPS][2][KP_MAX];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = blockIdx.x * AGG_WARPS + w;
    const int c0 = blockIdx.y * (CJ * 32);                  // channel slab of this block
    if (n >= nq) return;

    // kernel point of this lane (lanes 0..K-1 and 16..16+K-1 serve two neighbours per iteration)
    const int kl = lane & 15;
    float kx = 0.f, ky = 0.f, kz = 0.f;
    if (kl < K) { kx = kpts[3 * kl]; ky = kpts[3 * kl + 1]; kz = kpts[3 * kl + 2]; }
    const float qx = q_pts[3 * (size_t)n], qy = q_pts[3 * (size_t)n + 1], qz = q_pts[3 * (size_t)n + 2];

    float acc[KP_MAX - 1][CJ];
#pragma unroll
    for (int k = 0; k < KP_MAX - 1; k++)
#pragma unroll
        for (int j = 0; j < CJ; j++) acc[k][j] = 0.f;
    int cnt = 0;

    const IdxT* row = idx + (size_t)n * idx_stride;
    for (int h0 = 0; h0 < H; h0 += 2) {
        // lanes 0..15 -> neighbour h0, lanes 16..31 -> neighbour h0+1
        const int hh = h0 + (lane >> 4);
        long long jn = hh < H ? (long long)row[hh] : (long long)ns;
        const bool valid = jn >= 0 && jn < ns;
        float wgt = 0.f;
        if (valid && kl < K) {
            float dx = s_pts[3 * (size_t)jn] - qx - kx, dy = s_pts[3 * (size_t)jn + 1] - qy - ky,
                  dz = s_pts[3 * (size_t)jn + 2] - qz - kz;
            wgt = fmaxf(0.f, 1.f - sqrtf(dx * dx + dy * dy + dz * dz) * inv_extent);
        }
        s_w[w][lane >> 4][kl] = wgt;
        const long long j0 = __shfl_sync(0xffffffffu, jn, 0), j1 = __shfl_sync(0xffffffffu, jn, 16);
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const long long j = t == 0 ? j0 : j1;
            if (j < 0 || j >= ns) continue;            // shadow neighbour: zero weights, zero features
            if (blockIdx.y == 0) cnt += rowflag[j];
            float f[CJ];
#pragma unroll
            for (int jj = 0; jj < CJ; jj++) {
                int c = c0 + jj * 32 + lane;
                f[jj] = c < cin ? __ldg(x + (size_t)j * ldx + c) : 0.f;
            }
            const float4* wv = reinterpret_cast<const float4*>(s_w[w][t]);
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) {
                float4 ww = wv[k4];
                float wk[4] = { ww.x, ww.y, ww.z, ww.w };
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    int k = k4 * 4 + u;
                    if (k < KP_MAX - 1) {
#pragma unroll
                        for (int jj = 0; jj < CJ; jj++) acc[k][jj] = fmaf(wk[u], f[jj], acc[k][jj]);
                    }
                }
            }
        }
        __syncwarp();
    }
    // wf[n][k*cin + c]   (fp32, or split into bf16 hi/lo for the tensor-core contraction)
#pragma unroll
    for (int k = 0; k < KP_MAX - 1; k++) {
        if (k < K) {
#pragma unroll
            for (int jj = 0; jj < CJ; jj++) {
                int c = c0 + jj * 32 + lane;
                if (c < cin) {
                    if (SPLIT) {
                        __nv_bfloat16 h = __float2bfloat16_rn(acc[k][jj]);
                        size_t e = (size_t)n * ldk + (size_t)k * cin + c;
                        wf_hi[e] = h;
                        wf_lo[e] = __float2bfloat16_rn(acc[k][jj] - __bfloat162float(h));
                    } else {
                        wf[(size_t)n * ldk + (size_t)k * cin + c] = acc[k][jj];
                    }
                }
            }
        }
    }
    if (blockIdx.y == 0 && lane == 0) inv_cnt[n] = 1.0f / (float)(cnt > 1 ? cnt : 1);
}

template <typename IdxT, bool SPLIT>
static int launch_agg(const float* q_pts, int nq, const float* s_pts, int ns, const IdxT* idx, int H, int idx_stride, const float* x,
                      int cin, int ldx, const uint8_t* rowflag, const float* kpts, int K, float inv_extent, float* wf,
                      __nv_bfloat16* wf_hi, __nv_bfloat16* wf_lo, int ldk, float* inv_cnt, cudaStream_t st)
{
    dim3 block(AGG_WARPS * 32);
    unsigned gx = (unsigned)cdiv64(nq, AGG_WARPS);
    if (cin <= 32) {
        k_kpconv_aggregate<IdxT, 1, SPLIT><<<dim3(gx, 1), block, 0, st>>>(q_pts, nq, s_pts, ns, idx, H, idx_stride, x, cin, ldx, rowflag, kpts, K, inv_extent, wf, wf_hi, wf_lo, ldk, inv_cnt);
    } else if (cin <= 64) {
        k_kpconv_aggregate<IdxT, 2, SPLIT><<<dim3(gx, 1), block, 0, st>>>(q_pts, nq, s_pts, ns, idx, H, idx_stride, x, cin, ldx, rowflag, kpts, K, inv_extent, wf, wf_hi, wf_lo, ldk, inv_cnt);
    } else {
        unsigned gy = (unsigned)cdiv64(cin, 128);
        k_kpconv_aggregate<IdxT, 4, SPLIT><<<dim3(gx, gy), block, 0, st>>>(q_pts, nq, s_pts, ns, idx, H, idx_stride, x, cin, ldx, rowflag, kpts, K, inv_extent, wf, wf_hi, wf_lo, ldk, inv_cnt);
    }
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int gemm_dev(const float* A, int lda, const float* B, int ldb, int b_is_nk, float* C, int ldc, int M, int N, int K,
             const float* row_scale, cudaStream_t st);   // gemm.cu
bool gemm_tc_shape_ok(int M, int N, int K);             // gemm_tc.cu
int gemm_tc_presplit_dev(const void* a_hi, const void* a_lo, int ldk, const float* B, int ldb, int b_is_nk, float* C, int ldc, int M, int N,
                         int K, const float* row_scale, cudaStream_t st);
int gemm_force_simt_get();

size_t kpconv_ws_bytes(int64_t nq, int64_t ns, int32_t cin, int32_t K)
{
    size_t ldk = ((size_t)K * cin + 7) / 8 * 8;
    return align_up((size_t)nq * ldk * sizeof(float), 256) + align_up((size_t)nq * sizeof(float), 256) + align_up((size_t)ns, 256) + 1024;
}

// weights: [K, cin, cout] row-major (the reference's Parameter layout)
int kpconv_forward_dev(const float* q_pts, int64_t nq, const float* s_pts, int64_t ns, const void* idx, int idx_is_i64, int32_t H,
                       int32_t idx_stride, const float* x, int32_t cin, const float* kpts, int32_t K, float kp_extent,
                       const float* weights, int32_t cout, float* out, void* ws, size_t ws_bytes, cudaStream_t st)
{
    PCRCG_REQUIRE(K >= 1 && K <= KP_MAX - 1, "kpconv: kernel_size must be in [1,15]");
    PCRCG_REQUIRE(cin >= 1 && cout >= 1 && H >= 0 && idx_stride >= H, "kpconv: bad dimensions");
    PCRCG_REQUIRE(nq < (1ll << 31) && ns < (1ll << 31), "kpconv: too many points");
    PCRCG_REQUIRE(kp_extent > 0.f, "kpconv: KP_extent must be positive");
    if (nq == 0) return PCRCG_OK;
    const int KC = K * cin;
    const bool tc = !gemm_force_simt_get() && gemm_tc_shape_ok((int)nq, cout, KC);
    const int ldk = tc ? (KC + 7) / 8 * 8 : KC;
    Workspace W(ws, ws_bytes);
    float* wf = W.take<float>((size_t)nq * ldk);
    float* inv_cnt = W.take<float>((size_t)nq);
    uint8_t* rowflag = W.take<uint8_t>((size_t)(ns > 0 ? ns : 1));
    PCRCG_REQUIRE(ws != nullptr && W.ok(), "kpconv: workspace too small (%zu < %zu)", ws_bytes, W.off);
    __nv_bfloat16* wf_hi = (__nv_bfloat16*)wf;
    __nv_bfloat16* wf_lo = wf_hi + (size_t)nq * ldk;
    const float inv_extent = 1.0f / kp_extent;
    {
        ProfScope prof(PC_KPCONV_AGG, st, 2);
        if (ns > 0) {
            k_row_positive<<<(unsigned)cdiv64(ns, 8), 256, 0, st>>>(x, (int)ns, cin, cin, rowflag);
            PCRCG_CUDA(cudaGetLastError());
        }
        int rc;
        if (idx_is_i64) {
            rc = tc ? launch_agg<long long, true>(q_pts, (int)nq, s_pts, (int)ns, (const long long*)idx, H, idx_stride, x, cin, cin, rowflag, kpts, K, inv_extent, nullptr, wf_hi, wf_lo, ldk, inv_cnt, st)
                    : launch_agg<long long, false>(q_pts, (int)nq, s_pts, (int)ns, (const long long*)idx, H, idx_stride, x, cin, cin, rowflag, kpts, K, inv_extent, wf, nullptr, nullptr, ldk, inv_cnt, st);
        } else {
            rc = tc ? launch_agg<int, true>(q_pts, (int)nq, s_pts, (int)ns, (const int*)idx, H, idx_stride, x, cin, cin, rowflag, kpts, K, inv_extent, nullptr, wf_hi, wf_lo, ldk, inv_cnt, st)
                    : launch_agg<int, false>(q_pts, (int)nq, s_pts, (int)ns, (const int*)idx, H, idx_stride, x, cin, cin, rowflag, kpts, K, inv_extent, wf, nullptr, nullptr, ldk, inv_cnt, st);
        }
        if (rc) return rc;
    }
    if (tc) {
        ProfScope prof(PC_GEMM, st, 1);
        return gemm_tc_presplit_dev(wf_hi, wf_lo, ldk, weights, cout, 0, out, cout, (int)nq, cout, KC, inv_cnt, st);
    }
    return gemm_dev(wf, KC, weights, cout, 0, out, cout, (int)nq, cout, KC, inv_cnt, st);
}

}  // namespace pcrcg
