// Per-pair InstanceNorm (+LeakyReLU, + residual), max_pool / closest_pool gathers -- sm_100a.
//
// Replaces models/blocks.py:
//   BatchNormBlock.forward :456-465 (= InstanceNorm1d over the N rows of a fragment pair: biased
//        variance, eps 1e-5, no affine, no running stats), LeakyReLU(0.1) of :501,590,662,678
//   max_pool :86-102, closest_pool :71-83
// A "segment" is one normalisation group: the rows of one fragment pair inside a stacked batch
// (the reference processes one pair per batch, i.e. one segment).
#include "common.cuh"

namespace pcrcg {

// ---- column statistics: mean / rstd per (segment, column) ----------------------------------------
// grid (ceil(C/32), nseg), block 32 x 8 : two passes over the segment's rows (L2 resident).
__global__ void __launch_bounds__(256) k_colstats(const float* __restrict__ x, int ldx, int C, const int32_t* __restrict__ seg_starts,
                                                  float eps, float* __restrict__ mean, float* __restrict__ rstd)
{
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx, seg = blockIdx.y;
    const int r0 = seg_starts[seg], r1 = seg_starts[seg + 1];
    const bool ok = c < C;
    float s = 0.f;
    for (int r = r0 + ty; r < r1; r += 8) s += ok ? x[(size_t)r * ldx + c] : 0.f;
    red[ty][tx] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) tot += red[k][tx];
    const float n = (float)(r1 - r0);
    const float mu = n > 0.f ? tot / n : 0.f;
    __syncthreads();
    float v = 0.f;
    for (int r = r0 + ty; r < r1; r += 8) {
        float d = ok ? x[(size_t)r * ldx + c] - mu : 0.f;
        v = fmaf(d, d, v);
    }
    red[ty][tx] = v;
    __syncthreads();
    if (ty == 0 && ok) {
        float var = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) var += red[k][tx];
        var = n > 0.f ? var / n : 0.f;
        mean[(size_t)seg * C + c] = mu;
        rstd[(size_t)seg * C + c] = rsqrtf(var + eps);
    }
}

// out = act( (x - mean)*rstd  [ + (sc - sc_mean)*sc_rstd | + sc ] ),  act = LeakyReLU(slope) if slope >= 0
__global__ void __launch_bounds__(256) k_norm_act(const float* __restrict__ x, int ldx, int n, int C,
                                                  const int32_t* __restrict__ seg_starts, int nseg, const float* __restrict__ mean,
                                                  const float* __restrict__ rstd, const float* __restrict__ sc, int ldsc,
                                                  const float* __restrict__ sc_mean, const float* __restrict__ sc_rstd, float slope,
                                                  float* __restrict__ out, int ldo)
{
    const int c4 = (C + 3) >> 2;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * c4) return;
    const int r = (int)(e / c4), cb = (int)(e - (long long)r * c4) * 4;
    const int seg = nseg > 1 ? cloud_of(seg_starts, nseg, r) : 0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        int c = cb + u;
        if (c >= C) break;
        float v = x[(size_t)r * ldx + c];
        if (mean) v = (v - mean[(size_t)seg * C + c]) * rstd[(size_t)seg * C + c];
        if (sc) {
            float s = sc[(size_t)r * ldsc + c];
            if (sc_mean) s = (s - sc_mean[(size_t)seg * C + c]) * sc_rstd[(size_t)seg * C + c];
            v += s;
        }
        if (slope >= 0.f) v = v > 0.f ? v : v * slope;
        out[(size_t)r * ldo + c] = v;
    }
}

// ---- gathers --------------------------------------------------------------------------------------
// max over the listed rows; a shadow index contributes the zero row (models/blocks.py:95-101)
template <typename IdxT>
__global__ void __launch_bounds__(256) k_max_pool(const float* __restrict__ x, int ns, int C, int ldx, const IdxT* __restrict__ idx,
                                                  int nq, int H, int idx_stride, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= nq) return;
    const IdxT* row = idx + (size_t)n * idx_stride;
    for (int c = lane; c < C; c += 32) {
        float m = -INFINITY;
        for (int h = 0; h < H; h++) {
            long long j = (long long)row[h];
            float v = (j >= 0 && j < ns) ? __ldg(x + (size_t)j * ldx + c) : 0.f;
            m = fmaxf(m, v);
        }
        out[(size_t)n * C + c] = H > 0 ? m : 0.f;
    }
}

// x_pad[idx[n,0]]   (models/blocks.py:71-83)
template <typename IdxT>
__global__ void __launch_bounds__(256) k_closest_pool(const float* __restrict__ x, int ns, int C, int ldx, const IdxT* __restrict__ idx,
                                                      int nq, int idx_stride, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= nq) return;
    long long j = (long long)idx[(size_t)n * idx_stride];
    const bool ok = j >= 0 && j < ns;
    for (int c = lane; c < C; c += 32) out[(size_t)n * C + c] = ok ? __ldg(x + (size_t)j * ldx + c) : 0.f;
}

// ---- host side ---------------------------------------------------------------------------------
int colstats_dev(const float* x, int64_t n, int32_t C, const int32_t* seg_starts, int32_t nseg, float eps, float* mean, float* rstd,
                 cudaStream_t st)
{
    (void)n;
    PCRCG_REQUIRE(C >= 1 && nseg >= 1 && nseg < 65536, "instance norm: bad dimensions");
    ProfScope prof(PC_NORM, st, 1);
    k_colstats<<<dim3((unsigned)cdiv64(C, 32), (unsigned)nseg), 256, 0, st>>>(x, C, C, seg_starts, eps, mean, rstd);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int norm_act_dev(const float* x, int64_t n, int32_t C, const int32_t* seg_starts, int32_t nseg, const float* mean, const float* rstd,
                 const float* sc, const float* sc_mean, const float* sc_rstd, float slope, float* out, cudaStream_t st)
{
    if (n == 0) return PCRCG_OK;
    ProfScope prof(PC_NORM, st, 1);
    long long tot = (long long)n * ((C + 3) / 4);
    k_norm_act<<<(unsigned)cdiv64(tot, 256), 256, 0, st>>>(x, C, (int)n, C, seg_starts, nseg, mean, rstd, sc, C, sc_mean, sc_rstd, slope,
                                                          out, C);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int max_pool_dev(const float* x, int64_t ns, int32_t C, const void* idx, int idx_is_i64, int64_t nq, int32_t H, int32_t idx_stride,
                 float* out, cudaStream_t st)
{
    if (nq == 0) return PCRCG_OK;
    ProfScope prof(PC_POOL, st, 1);
    unsigned g = (unsigned)cdiv64(nq, 8);
    if (idx_is_i64) k_max_pool<long long><<<g, 256, 0, st>>>(x, (int)ns, C, C, (const long long*)idx, (int)nq, H, idx_stride, out);
    else k_max_pool<int><<<g, 256, 0, st>>>(x, (int)ns, C, C, (const int*)idx, (int)nq, H, idx_stride, out);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int closest_pool_dev(const float* x, int64_t ns, int32_t C, const void* idx, int idx_is_i64, int64_t nq, int32_t idx_stride, float* out,
                     cudaStream_t st)
{
    if (nq == 0) return PCRCG_OK;
    ProfScope prof(PC_POOL, st, 1);
    unsigned g = (unsigned)cdiv64(nq, 8);
    if (idx_is_i64) k_closest_pool<long long><<<g, 256, 0, st>>>(x, (int)ns, C, C, (const long long*)idx, (int)nq, idx_stride, out);
    else k_closest_pool<int><<<g, 256, 0, st>>>(x, (int)ns, C, C, (const int*)idx, (int)nq, idx_stride, out);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

}  // namespace pcrcg
