// Per-pair InstanceNorm (+LeakyReLU, + residual), max_pool / closest_pool gathers -- sm_100a.
//
// Replaces models/blocks.py:
//   BatchNormBlock.forward :456-465 (= InstanceNorm1d over the N rows of a fragment pair: biased
//        variance, eps 1e-5, no affine, no running stats), LeakyReLU(0.1) of :501,590,662,678
//   max_pool :86-102, closest_pool :71-83
// A "segment" is one normalisation group: the rows of one fragment pair inside a stacked batch
// (the reference processes one pair per batch, i.e. one segment).
#include "common.cuh"

#include <cuda_bf16.h>

namespace pcrcg {

// ---- column statistics: mean / rstd per (segment, column) ----------------------------------------
// Pass 1: each block reduces CS_ROWS consecutive rows x 32 columns (one coalesced read of x) and adds
// its partial sum / sum of squares in fp64 to the segment's accumulators; a block that straddles a
// segment boundary flushes once per segment.  Pass 2 turns them into mean and rsqrt(var + eps)
// (biased variance, E[x^2] - mean^2 evaluated in fp64).
constexpr int CS_ROWS = 512;

__global__ void __launch_bounds__(256) k_colstats_partial(const float* __restrict__ x, int ldx, int n, int C,
                                                          const int32_t* __restrict__ seg_starts, int nseg, double* __restrict__ acc)
{
    __shared__ double red[2][8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const bool ok = c < C;
    int r = blockIdx.y * CS_ROWS;
    const int rend = min(n, r + CS_ROWS);
    int seg = nseg > 1 ? cloud_of(seg_starts, nseg, r) : 0;
    while (r < rend) {
        const int stop = min(rend, seg_starts[seg + 1]);
        float s = 0.f, s2 = 0.f;                      // <= 64 rows per thread: fp32 partials are safe
        for (int i = r + ty; i < stop; i += 8) {
            float v = ok ? x[(size_t)i * ldx + c] : 0.f;
            s += v;
            s2 = fmaf(v, v, s2);
        }
        red[0][ty][tx] = (double)s;
        red[1][ty][tx] = (double)s2;
        __syncthreads();
        if (ty < 2 && ok) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) t += red[ty][k][tx];
            atomicAdd(acc + ((size_t)seg * 2 + ty) * C + c, t);
        }
        __syncthreads();
        r = stop;
        seg++;
    }
}

__global__ void __launch_bounds__(256) k_colstats_final(const double* __restrict__ acc, const int32_t* __restrict__ seg_starts, int nseg,
                                                        int C, float eps, float* __restrict__ mean, float* __restrict__ rstd)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nseg * C) return;
    const int seg = e / C, c = e - seg * C;
    const double n = (double)(seg_starts[seg + 1] - seg_starts[seg]);
    double mu = 0.0, var = 0.0;
    if (n > 0) {
        mu = acc[((size_t)seg * 2 + 0) * C + c] / n;
        var = acc[((size_t)seg * 2 + 1) * C + c] / n - mu * mu;
        if (var < 0.0) var = 0.0;
    }
    mean[e] = (float)mu;
    rstd[e] = (float)(1.0 / sqrt(var + (double)eps));
}

// out = act( (x - mean)*rstd  [ + (sc - sc_mean)*sc_rstd | + sc ] ),  act = LeakyReLU(slope) if slope >= 0
__global__ void __launch_bounds__(256) k_norm_act(const float* __restrict__ x, int ldx, int n, int C,
                                                  const int32_t* __restrict__ seg_starts, int nseg, const float* __restrict__ mean,
                                                  const float* __restrict__ rstd, const float* __restrict__ sc, int ldsc,
                                                  const float* __restrict__ sc_mean, const float* __restrict__ sc_rstd, float slope,
                                                  float* __restrict__ out, int ldo, __nv_bfloat16* __restrict__ hi,
                                                  __nv_bfloat16* __restrict__ lo, int lds)
{
    const int c4 = (C + 3) >> 2;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * c4) return;
    const int r = (int)(e / c4), cb = (int)(e - (long long)r * c4) * 4;
    const int seg = nseg > 1 ? cloud_of(seg_starts, nseg, r) : 0;
    float vv[4] = { 0.f, 0.f, 0.f, 0.f };
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
        if (out != nullptr) out[(size_t)r * ldo + c] = v;
        vv[u] = v;
    }
    if (hi != nullptr) {          // bf16 (hi, lo) planes of the result for the next tensor-core contraction (lds % 8 == 0)
        __align__(8) __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            h[u] = __float2bfloat16_rn(vv[u]);
            l[u] = __float2bfloat16_rn(vv[u] - __bfloat162float(h[u]));
        }
        *reinterpret_cast<uint2*>(hi + (size_t)r * lds + cb) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(lo + (size_t)r * lds + cb) = *reinterpret_cast<const uint2*>(l);
    }
}

// Vectorised variant (C % 4 == 0): a thread owns one 4-channel group (mean / rstd in registers) and
// NA_UNROLL rows whose float4 loads are all issued before the arithmetic; a block covers
// (256 / column groups) * NA_UNROLL consecutive rows and re-reads the statistics if it crosses a segment.

__device__ __forceinline__ float4 na_apply(float4 v, float4 mu, float4 rs)
{
    return make_float4((v.x - mu.x) * rs.x, (v.y - mu.y) * rs.y, (v.z - mu.z) * rs.z, (v.w - mu.w) * rs.w);
}

template <int NA_UNROLL, int MINB>
__global__ void __launch_bounds__(256, MINB) k_norm_act_v4(const float* __restrict__ x, int n, int C, const int32_t* __restrict__ seg_starts,
                                                     int nseg, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ sc, const float* __restrict__ sc_mean,
                                                     const float* __restrict__ sc_rstd, float slope, float* __restrict__ out,
                                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int lds,
                                                     int tcols, int trows, uint8_t* __restrict__ rowflag,
                                                     const __nv_bfloat16* __restrict__ sc_hi, const __nv_bfloat16* __restrict__ sc_lo, int sc_lds)
{
    const int c4 = C >> 2;
    float rsum[NA_UNROLL];
#pragma unroll
    for (int u = 0; u < NA_UNROLL; u++) rsum[u] = 0.f;
    const int tx = threadIdx.x % tcols, ty = threadIdx.x / tcols;
    if (ty >= trows) return;
    const int R0 = blockIdx.x * (trows * NA_UNROLL);
    for (int cg = tx; cg < c4; cg += tcols) {
        float4 v[NA_UNROLL], s[NA_UNROLL];
        int rr[NA_UNROLL];
#pragma unroll
        for (int u = 0; u < NA_UNROLL; u++) {
            rr[u] = R0 + ty + u * trows;
            if (rr[u] < n) {
                v[u] = *reinterpret_cast<const float4*>(x + (size_t)rr[u] * C + 4 * cg);
                if (sc) s[u] = *reinterpret_cast<const float4*>(sc + (size_t)rr[u] * C + 4 * cg);
                else if (sc_hi) {
                    // the shortcut exists only as bf16 (hi, lo) planes (a block output whose fp32 copy was never written)
                    const uint2 a = *reinterpret_cast<const uint2*>(sc_hi + (size_t)rr[u] * sc_lds + 4 * cg);
                    const uint2 b = *reinterpret_cast<const uint2*>(sc_lo + (size_t)rr[u] * sc_lds + 4 * cg);
                    s[u] = make_float4(__uint_as_float(a.x << 16) + __uint_as_float(b.x << 16), __uint_as_float(a.x & 0xffff0000u) + __uint_as_float(b.x & 0xffff0000u),
                                       __uint_as_float(a.y << 16) + __uint_as_float(b.y << 16), __uint_as_float(a.y & 0xffff0000u) + __uint_as_float(b.y & 0xffff0000u));
                }
            }
        }
        int seg = -1;
        float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = make_float4(1.f, 1.f, 1.f, 1.f), smu = mu, srs = rs;
#pragma unroll
        for (int u = 0; u < NA_UNROLL; u++) {
            const int r = rr[u];
            if (r >= n) break;
            if (seg < 0 || r >= seg_starts[seg + 1]) {
                seg = nseg > 1 ? cloud_of(seg_starts, nseg, r) : 0;
                if (mean) {
                    mu = *reinterpret_cast<const float4*>(mean + (size_t)seg * C + 4 * cg);
                    rs = *reinterpret_cast<const float4*>(rstd + (size_t)seg * C + 4 * cg);
                }
                if (sc_mean) {
                    smu = *reinterpret_cast<const float4*>(sc_mean + (size_t)seg * C + 4 * cg);
                    srs = *reinterpret_cast<const float4*>(sc_rstd + (size_t)seg * C + 4 * cg);
                }
            }
            float4 o = na_apply(v[u], mu, rs);
            if (sc || sc_hi) {
                const float4 t = na_apply(s[u], smu, srs);
                o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
            }
            if (slope >= 0.f) {
                o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
                o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
            }
            if (out != nullptr) *reinterpret_cast<float4*>(out + (size_t)r * C + 4 * cg) = o;
            rsum[u] += (o.x + o.y) + (o.z + o.w);
            if (hi != nullptr) {
                const float vv[4] = { o.x, o.y, o.z, o.w };
                __align__(8) __nv_bfloat16 h[4], l[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    h[k] = __float2bfloat16_rn(vv[k]);
                    l[k] = __float2bfloat16_rn(vv[k] - __bfloat162float(h[k]));
                }
                *reinterpret_cast<uint2*>(hi + (size_t)r * lds + 4 * cg) = *reinterpret_cast<const uint2*>(h);
                *reinterpret_cast<uint2*>(lo + (size_t)r * lds + 4 * cg) = *reinterpret_cast<const uint2*>(l);
            }
        }
    }
    // (row sum > 0) flags for the KPConv neighbour count (models/blocks.py:369-370): a row lives in tcols <= 32
    // consecutive lanes of one warp in this mode, so a shuffle reduction finishes the sum deterministically
    if (rowflag != nullptr) {
#pragma unroll
        for (int u = 0; u < NA_UNROLL; u++) {
            float v = rsum[u];
            for (int o = tcols >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const int r = R0 + ty + u * trows;
            if (tx == 0 && r < n) rowflag[r] = v > 0.f ? 1 : 0;
        }
    }
}

// ---- gathers --------------------------------------------------------------------------------------
// max over the listed rows; a shadow index contributes the zero row (models/blocks.py:95-101).
// One warp per output row; the row's indices are loaded once (coalesced) and broadcast by shuffle.
template <typename IdxT>
__global__ void __launch_bounds__(256) k_max_pool(const float* __restrict__ x, int ns, int C, int ldx, const IdxT* __restrict__ idx,
                                                  int nq, int H, int idx_stride, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= nq) return;
    const IdxT* row = idx + (size_t)n * idx_stride;
    if ((C & 3) == 0) {
        for (int cb = 0; cb < C; cb += 128) {               // warp-uniform trip counts (shuffles inside)
            const int c = cb + 4 * lane;
            const bool act = c < C;
            float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            for (int h0 = 0; h0 < H; h0 += 32) {            // 32 indices per round: one coalesced load, broadcast by shuffle
                int jl = ns;
                if (h0 + lane < H) { long long v = (long long)row[h0 + lane]; jl = (v >= 0 && v < ns) ? (int)v : ns; }
                const int hn = min(32, H - h0);
#pragma unroll 4
                for (int h = 0; h < hn; h++) {
                    const int j = __shfl_sync(0xffffffffu, jl, h);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (act && j < ns) v = __ldg(reinterpret_cast<const float4*>(x + (size_t)j * ldx + c));
                    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
                }
            }
            if (H == 0) m = make_float4(0.f, 0.f, 0.f, 0.f);
            if (act) *reinterpret_cast<float4*>(out + (size_t)n * C + c) = m;
        }
        return;
    }
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

// max_pool of features that exist as bf16 (hi, lo) planes: the value of an entry is hi + lo (exact in fp32); the winner's
// (hi, lo) pair is copied, so the output planes represent the maximum exactly; a shadow index contributes (0, 0).
template <typename IdxT>
__global__ void __launch_bounds__(256) k_max_pool_planes(const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo, int ns,
                                                         int C, int ldx, const IdxT* __restrict__ idx, int nq, int H, int idx_stride,
                                                         __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_lo, int ldo)
{
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= nq) return;
    const IdxT* row = idx + (size_t)n * idx_stride;
    for (int cb = 0; cb < C; cb += 128) {                   // warp-uniform trip counts (shuffles inside)
        const int c = cb + 4 * lane;
        const bool act = c < C;
        float m[4] = { -INFINITY, -INFINITY, -INFINITY, -INFINITY };
        uint32_t bh[4] = { 0, 0, 0, 0 }, bl[4] = { 0, 0, 0, 0 };          // winner's hi / lo patterns (upper 16 bits)
        for (int h0 = 0; h0 < H; h0 += 32) {
            int jl = ns;
            if (h0 + lane < H) { long long v = (long long)row[h0 + lane]; jl = (v >= 0 && v < ns) ? (int)v : ns; }
            const int hn = min(32, H - h0);
#pragma unroll 4
            for (int h = 0; h < hn; h++) {
                const int j = __shfl_sync(0xffffffffu, jl, h);
                uint2 a = make_uint2(0u, 0u), b = make_uint2(0u, 0u);
                if (act && j < ns) {
                    a = __ldg(reinterpret_cast<const uint2*>(x_hi + (size_t)j * ldx + c));
                    b = __ldg(reinterpret_cast<const uint2*>(x_lo + (size_t)j * ldx + c));
                }
                const uint32_t ah[4] = { a.x << 16, a.x & 0xffff0000u, a.y << 16, a.y & 0xffff0000u };
                const uint32_t al[4] = { b.x << 16, b.x & 0xffff0000u, b.y << 16, b.y & 0xffff0000u };
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float v = __uint_as_float(ah[k]) + __uint_as_float(al[k]);
                    if (v > m[k]) { m[k] = v; bh[k] = ah[k]; bl[k] = al[k]; }
                }
            }
        }
        if (act) {
            *reinterpret_cast<uint2*>(o_hi + (size_t)n * ldo + c) = make_uint2((bh[0] >> 16) | bh[1], (bh[2] >> 16) | bh[3]);
            *reinterpret_cast<uint2*>(o_lo + (size_t)n * ldo + c) = make_uint2((bl[0] >> 16) | bl[1], (bl[2] >> 16) | bl[3]);
        }
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

// Descriptor head of KPFCNN.forward (models/architectures.py:572-582): L2-normalised features (F.normalize, eps 1e-12)
// and the two sigmoid scores, clamped to [0,1], NaN / Inf -> 0 (regular_score, :176-179).  One warp per row.
__global__ void __launch_bounds__(256) k_descriptor_head(const float* __restrict__ x, int n, int F, float* __restrict__ feats,
                                                         float* __restrict__ overlap, float* __restrict__ saliency)
{
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    const float* row = x + (size_t)r * (F + 2);
    float ss = 0.f;
    for (int c = lane; c < F; c += 32) { float v = row[c]; ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    for (int c = lane; c < F; c += 32) feats[(size_t)r * F + c] = row[c] * inv;
    if (lane < 2) {
        float v = 1.0f / (1.0f + expf(-row[F + lane]));
        v = fminf(fmaxf(v, 0.f), 1.f);
        if (!(v == v) || isinf(v)) v = 0.f;
        (lane == 0 ? overlap : saliency)[r] = v;
    }
}

static int g_norm_v4 = 1, g_norm_variant = 0;
void dense_set_norm_v4(int v) { g_norm_v4 = v; }
void dense_set_norm_variant(int v) { g_norm_variant = v; }

// ---- host side ---------------------------------------------------------------------------------
int colstats_dev(const float* x, int64_t n, int32_t C, const int32_t* seg_starts, int32_t nseg, float eps, float* mean, float* rstd,
                 cudaStream_t st)
{
    PCRCG_REQUIRE(C >= 1 && nseg >= 1 && nseg < 65536 && n < (1ll << 31), "instance norm: bad dimensions");
    PCRCG_TRY(pool_setup());
    ProfScope prof(PC_NORM, st, 2);
    double* acc = nullptr;
    const size_t acc_bytes = (size_t)nseg * 2 * C * sizeof(double);
    PCRCG_CUDA(cudaMallocAsync((void**)&acc, acc_bytes, st));
    PCRCG_CUDA(cudaMemsetAsync(acc, 0, acc_bytes, st));
    if (n > 0) k_colstats_partial<<<dim3((unsigned)cdiv64(C, 32), (unsigned)cdiv64(n, CS_ROWS)), 256, 0, st>>>(x, C, (int)n, C, seg_starts, nseg, acc);
    k_colstats_final<<<(unsigned)cdiv64((int64_t)nseg * C, 256), 256, 0, st>>>(acc, seg_starts, nseg, C, eps, mean, rstd);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(acc, st);
    PCRCG_CUDA(e);
    return PCRCG_OK;
}

// acc [nseg][2][C] fp64 (sum, sum of squares) as accumulated by the contraction epilogue (gemm_tc.cu) -> mean / rstd
int colstats_final_dev(const double* acc, const int32_t* seg_starts, int32_t nseg, int32_t C, float eps, float* mean, float* rstd,
                       cudaStream_t st)
{
    PCRCG_REQUIRE(C >= 1 && nseg >= 1 && nseg < 65536, "instance norm: bad dimensions");
    ProfScope prof(PC_NORM, st, 1);
    k_colstats_final<<<(unsigned)cdiv64((int64_t)nseg * C, 256), 256, 0, st>>>(acc, seg_starts, nseg, C, eps, mean, rstd);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int norm_act_planes_dev(const float* x, int64_t n, int32_t C, const int32_t* seg_starts, int32_t nseg, const float* mean, const float* rstd,
                        const float* sc, const float* sc_mean, const float* sc_rstd, float slope, float* out, void* split_hi, void* split_lo,
                        int32_t split_ld, uint8_t* rowflag, const void* sc_hi, const void* sc_lo, int32_t sc_ld, cudaStream_t st);

int norm_act_dev(const float* x, int64_t n, int32_t C, const int32_t* seg_starts, int32_t nseg, const float* mean, const float* rstd,
                 const float* sc, const float* sc_mean, const float* sc_rstd, float slope, float* out, void* split_hi, void* split_lo,
                 int32_t split_ld, uint8_t* rowflag, cudaStream_t st)
{
    return norm_act_planes_dev(x, n, C, seg_starts, nseg, mean, rstd, sc, sc_mean, sc_rstd, slope, out, split_hi, split_lo, split_ld, rowflag,
                               nullptr, nullptr, 0, st);
}

// sc_hi / sc_lo (optional, instead of sc): the shortcut as bf16 (hi, lo) planes, row pitch sc_ld
int norm_act_planes_dev(const float* x, int64_t n, int32_t C, const int32_t* seg_starts, int32_t nseg, const float* mean, const float* rstd,
                        const float* sc, const float* sc_mean, const float* sc_rstd, float slope, float* out, void* split_hi, void* split_lo,
                        int32_t split_ld, uint8_t* rowflag, const void* sc_hi, const void* sc_lo, int32_t sc_ld, cudaStream_t st)
{
    if (n == 0) return PCRCG_OK;
    PCRCG_REQUIRE(sc_hi == nullptr || (sc == nullptr && sc_lo != nullptr && sc_ld % 8 == 0 && sc_ld >= C && C % 4 == 0 && n < (1ll << 31) && g_norm_v4),
                  "norm_act: a plane shortcut needs C %% 4 == 0, both planes and no fp32 shortcut");
    PCRCG_REQUIRE(split_hi == nullptr || (split_lo != nullptr && split_ld % 8 == 0 && split_ld >= C), "norm_act: bad split geometry");
    PCRCG_REQUIRE(out != nullptr || split_hi != nullptr, "norm_act: no output requested (out and the split planes are both NULL)");
    ProfScope prof(PC_NORM, st, 1);
    if (C % 4 == 0 && n < (1ll << 31) && g_norm_v4) {
        const int c4 = C / 4;
        // power-of-two column-group counts only in flag mode (shuffle reduction inside tcols lanes)
        const bool flag_ok = rowflag != nullptr && (c4 & (c4 - 1)) == 0;
        const int tcols = flag_ok ? (c4 < 32 ? c4 : 32) : (c4 < 256 ? c4 : 256), trows = 256 / tcols;
        PCRCG_REQUIRE(rowflag == nullptr || flag_ok, "norm_act: row flags need a power-of-two channel count");
#define PCRCG_NA(U_, B_) k_norm_act_v4<U_, B_><<<(unsigned)cdiv64(n, trows * U_), 256, 0, st>>>( \
            x, (int)n, C, seg_starts, nseg, mean, rstd, sc, sc_mean, sc_rstd, slope, out, (__nv_bfloat16*)split_hi, (__nv_bfloat16*)split_lo, \
            split_ld, tcols, trows, rowflag, (const __nv_bfloat16*)sc_hi, (const __nv_bfloat16*)sc_lo, sc_ld)
        // launch shape measured per case on B200 (tools/bench_norm_apply.py): the row-flag mode (narrow tensors, extra shuffles)
        // prefers more resident warps, the plain streaming mode more loads in flight per thread
        switch (g_norm_variant > 0 ? g_norm_variant : (rowflag != nullptr ? 4 : 2)) {
            case 1: PCRCG_NA(4, 3); break;
            case 2: PCRCG_NA(8, 2); break;
            case 4: PCRCG_NA(2, 4); break;
            default: PCRCG_NA(4, 2); break;
        }
#undef PCRCG_NA
        PCRCG_CUDA(cudaGetLastError());
        return PCRCG_OK;
    }
    PCRCG_REQUIRE(rowflag == nullptr, "norm_act: row flags need C % 4 == 0");
    long long tot = (long long)n * ((C + 3) / 4);
    k_norm_act<<<(unsigned)cdiv64(tot, 256), 256, 0, st>>>(x, C, (int)n, C, seg_starts, nseg, mean, rstd, sc, C, sc_mean, sc_rstd, slope,
                                                          out, C, (__nv_bfloat16*)split_hi, (__nv_bfloat16*)split_lo, split_ld);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int descriptor_head_dev(const float* x, int64_t n, int32_t F, float* feats, float* overlap, float* saliency, cudaStream_t st)
{
    PCRCG_REQUIRE(F >= 1 && n >= 0 && n < (1ll << 31), "descriptor head: bad dimensions");
    if (n == 0) return PCRCG_OK;
    ProfScope prof(PC_NORM, st, 1);
    k_descriptor_head<<<(unsigned)cdiv64(n, 8), 256, 0, st>>>(x, (int)n, F, feats, overlap, saliency);
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

int max_pool_planes_dev(const void* x_hi, const void* x_lo, int64_t ns, int32_t C, int32_t ldx, const void* idx, int idx_is_i64, int64_t nq,
                        int32_t H, int32_t idx_stride, void* o_hi, void* o_lo, int32_t ldo, cudaStream_t st)
{
    PCRCG_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && ldx >= C && ldo >= C, "max_pool (planes): C and the row pitches must be multiples of 4");
    if (nq == 0) return PCRCG_OK;
    ProfScope prof(PC_POOL, st, 1);
    const unsigned g = (unsigned)cdiv64(nq, 8);
    if (idx_is_i64)
        k_max_pool_planes<long long><<<g, 256, 0, st>>>((const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, (int)ns, C, ldx, (const long long*)idx,
                                                        (int)nq, H, idx_stride, (__nv_bfloat16*)o_hi, (__nv_bfloat16*)o_lo, ldo);
    else
        k_max_pool_planes<int><<<g, 256, 0, st>>>((const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, (int)ns, C, ldx, (const int*)idx, (int)nq,
                                                  H, idx_stride, (__nv_bfloat16*)o_hi, (__nv_bfloat16*)o_lo, ldo);
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
