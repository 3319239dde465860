// KPConv forward (rigid kernel, linear influence, sum aggregation) -- sm_100a.
//
// Replaces models/blocks.py:229-374 (KPConv.forward, non-deformable branch).  Math:
//   w[n,k,h] = max(0, 1 - |s[idx[n,h]] - q[n] - kp[k]| / KP_extent)          (shadow neighbour -> 0)
//   wf[n,k,:] = sum_h w[n,k,h] * x[idx[n,h],:]
//   out[n,:]  = (sum_k wf[n,k,:] @ W[k]) / max(1, #{h : sum_c x[idx[n,h],c] > 0})
//
// Two stages: the AGGREGATION wf (this file) and the [Nq, K*Cin] x [K*Cin, Cout] CONTRACTION (gemm_tc.cu / gemm.cu) with the
// 1/count row scale -- and optionally the InstanceNorm statistics of the result -- in its epilogue.  Aggregation kernels, in
// dispatch order (kpconv_forward_dev / launch_agg):
//   k_kpconv_aggregate_bf16p   features given as bf16 (hi, lo) planes, cin % 64 == 0, H <= 64: persistent, software-pipelined,
//                              ldmatrix + mma.sync m16n8k16 bf16x3, register stores in kperm64 slab order   (every production layer)
//   k_kpconv_aggregate_bf16    same inputs, any H: one point per warp, 32-row staging, stmatrix transposed stores
//   k_kpconv_aggregate_small   cin <= 4 (first layer): one thread per (point, kernel point)
//   k_kpconv_small_fused       cin <= 4, whole KPConv in one kernel (opt-in, see g_small_fused)
//   k_kpconv_aggregate_mma64   fp32 features, cin % 64 == 0: cp.async staging + mma.sync m16n8k8 3xTF32
//   k_kpconv_aggregate_mma     fp32 features, cin % 8 == 0: direct gathers + 3xTF32
//   k_kpconv_aggregate         CUDA-core fallback for ragged cin (129: colour path) and the "aggregate_simt" parity anchor
#include "agg_ptx.cuh"

namespace pcrcg {

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

    // lanes 0..14 / 16..30 hold kernel point kl for the neighbour of their half-warp; lanes 15 / 31 are
    // spare and read the neighbour's "row sum > 0" flag.  (s - q) - kp is evaluated as s - (q + kp).
    const int kl = lane & 15, half = lane >> 4;
    float kx = 0.f, ky = 0.f, kz = 0.f;
    if (kl < K) {
        kx = kpts[3 * kl] + q_pts[3 * (size_t)n];
        ky = kpts[3 * kl + 1] + q_pts[3 * (size_t)n + 1];
        kz = kpts[3 * kl + 2] + q_pts[3 * (size_t)n + 2];
    }
    float acc[KP_MAX - 1][CJ];
#pragma unroll
    for (int k = 0; k < KP_MAX - 1; k++)
#pragma unroll
        for (int j = 0; j < CJ; j++) acc[k][j] = 0.f;
    int cnt = 0;

    const IdxT* row = idx + (size_t)n * idx_stride;
    const float* xc = x + c0 + lane;
    bool cok[CJ];
#pragma unroll
    for (int jj = 0; jj < CJ; jj++) cok[jj] = c0 + jj * 32 + lane < cin;
    float* sw = &s_w[w][half][kl];
    const float4* wv0 = reinterpret_cast<const float4*>(s_w[w][0]);
    const float4* wv1 = reinterpret_cast<const float4*>(s_w[w][1]);

    for (int h0 = 0; h0 < H; h0 += 2) {
        const int hh = h0 + half;
        int jn = ns;
        if (hh < H) { long long t = (long long)row[hh]; jn = (t >= 0 && t < ns) ? (int)t : ns; }
        const bool valid = jn < ns;
        float wgt = 0.f;
        bool pos = false;
        if (valid) {
            if (kl < K) {
                const float* sp = s_pts + 3 * (size_t)(unsigned)jn;
                const float dx = sp[0] - kx, dy = sp[1] - ky, dz = sp[2] - kz;
                const float d2 = fmaxf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)), 1e-30f);
                wgt = fmaxf(0.f, fmaf(-d2 * rsqrtf(d2), inv_extent, 1.f));
            } else if (kl == 15) {
                pos = rowflag[jn] != 0;
            }
        }
        *sw = wgt;
        cnt += __popc(__ballot_sync(0xffffffffu, pos));
        const int j0 = __shfl_sync(0xffffffffu, jn, 0), j1 = __shfl_sync(0xffffffffu, jn, 16);
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const int j = t == 0 ? j0 : j1;
            if (j >= ns) continue;                     // shadow neighbour: zero weights, zero features (warp-uniform)
            const float* xr = xc + (size_t)((unsigned)j * (unsigned)ldx);
            float f[CJ];
#pragma unroll
            for (int jj = 0; jj < CJ; jj++) f[jj] = cok[jj] ? __ldg(xr + jj * 32) : 0.f;
            const float4* wv = t == 0 ? wv0 : wv1;
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) {
                const float4 ww = wv[k4];
                const float wk[4] = { ww.x, ww.y, ww.z, ww.w };
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int k = k4 * 4 + u;
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
    const size_t obase = (size_t)n * ldk + c0 + lane;
#pragma unroll
    for (int k = 0; k < KP_MAX - 1; k++) {
        if (k < K) {
#pragma unroll
            for (int jj = 0; jj < CJ; jj++) {
                if (cok[jj]) {
                    const size_t e = obase + (size_t)k * cin + jj * 32;
                    if (SPLIT) {
                        const __nv_bfloat16 h = __float2bfloat16_rn(acc[k][jj]);
                        wf_hi[e] = h;
                        wf_lo[e] = __float2bfloat16_rn(acc[k][jj] - __bfloat162float(h));
                    } else {
                        wf[e] = acc[k][jj];
                    }
                }
            }
        }
    }
    if (blockIdx.y == 0 && lane == 0) inv_cnt[n] = 1.0f / (float)(cnt > 1 ? cnt : 1);
}

// ---------------------------------------------------------------------------------------------
// Tensor-core aggregation (cin % 8 == 0): per query point the [16 kernel points x H] influence matrix
// times the gathered [H x channels] feature rows is evaluated with warp-level mma.sync m16n8k8 TF32,
// operands split hi/lo ("3xTF32": hi*hi + hi*lo + lo*hi) so the result is fp32-accurate.
//   M = kernel point (15 padded to 16), K = neighbour (8 per step), N = 8 channels per tile.
//   Each lane computes exactly the 4 influence weights of its A fragment (kernel points g, g+8 x
//   neighbours t, t+4 of the step) -- every (kernel point, neighbour) pair is evaluated once per warp
//   and never leaves registers.  Channel c of tile nt, column g' is  c0 + g'*NT + nt, so a lane's B
//   fragment values for all NT tiles are NT CONSECUTIVE floats of the neighbour's feature row
//   (float4 gathers) and its D fragments cover 2*NT consecutive channels (vector stores).
constexpr uint32_t TF32_MASK = 0xffffe000u;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float influence(const float* __restrict__ sp, float kx, float ky, float kz, float inv_extent)
{
    const float dx = sp[0] - kx, dy = sp[1] - ky, dz = sp[2] - kz;
    const float d2 = fmaxf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)), 1e-30f);
    return fmaxf(0.f, fmaf(-d2 * rsqrtf(d2), inv_extent, 1.f));
}

constexpr int MMA_KS = 5;            // k-steps (8 neighbours each) whose weight fragments are held in registers at once

template <typename IdxT, int NT, bool SPLIT>
__global__ void __launch_bounds__(AGG_WARPS * 32) k_kpconv_aggregate_mma(
    const float* __restrict__ q_pts, int nq, const float* __restrict__ s_pts, int ns, const IdxT* __restrict__ idx, int H,
    int idx_stride, const float* __restrict__ x, int cin, int ldx, const uint8_t* __restrict__ rowflag,
    const float* __restrict__ kpts, int K, float inv_extent, float* __restrict__ wf, __nv_bfloat16* __restrict__ wf_hi,
    __nv_bfloat16* __restrict__ wf_lo, int ldk, float* __restrict__ inv_cnt)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = blockIdx.x * AGG_WARPS + w;
    if (n >= nq) return;
    const int g = lane >> 2, t = lane & 3;
    const int c0 = blockIdx.y * (NT * 8);

    const float qx = q_pts[3 * (size_t)n], qy = q_pts[3 * (size_t)n + 1], qz = q_pts[3 * (size_t)n + 2];
    const bool k1ok = g + 8 < K, k0ok = g < K;
    const int ka = k0ok ? g : 0, kb = k1ok ? g + 8 : 0;
    const float k0x = kpts[3 * ka] + qx, k0y = kpts[3 * ka + 1] + qy, k0z = kpts[3 * ka + 2] + qz;
    const float k1x = kpts[3 * kb] + qx, k1y = kpts[3 * kb + 1] + qy, k1z = kpts[3 * kb + 2] + qz;

    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; i++) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    int cnt = 0;
    const IdxT* row = idx + (size_t)n * idx_stride;
    const float* xl = x + c0 + g * NT;                    // this lane's NT consecutive channels

    for (int h0 = 0; h0 < H; h0 += MMA_KS * 8) {
        uint32_t ahi[MMA_KS][4], alo[MMA_KS][4];
        unsigned ja[MMA_KS], jb[MMA_KS];
        bool vas[MMA_KS], vbs[MMA_KS];
        // all index loads first (independent), then the flag loads of the counting lanes
#pragma unroll
        for (int s = 0; s < MMA_KS; s++) {
            const int ha = h0 + 8 * s + t, hb = ha + 4;
            long long ia = ha < H ? (long long)row[ha] : (long long)ns;
            long long ib = hb < H ? (long long)row[hb] : (long long)ns;
            vas[s] = ia >= 0 && ia < ns;
            vbs[s] = ib >= 0 && ib < ns;
            ja[s] = vas[s] ? (unsigned)ia : 0u;
            jb[s] = vbs[s] ? (unsigned)ib : 0u;
        }
        if (g == 0 && blockIdx.y == 0) {
            int fl[2 * MMA_KS];
#pragma unroll
            for (int s = 0; s < MMA_KS; s++) { fl[2 * s] = vas[s] ? (int)rowflag[ja[s]] : 0; fl[2 * s + 1] = vbs[s] ? (int)rowflag[jb[s]] : 0; }
#pragma unroll
            for (int s = 0; s < 2 * MMA_KS; s++) cnt += fl[s];
        }
#pragma unroll
        for (int s = 0; s < MMA_KS; s++) {
            const bool va = vas[s], vb = vbs[s];
            const float* spa = s_pts + 3 * (size_t)ja[s];
            const float* spb = s_pts + 3 * (size_t)jb[s];
            float w00 = (va && k0ok) ? influence(spa, k0x, k0y, k0z, inv_extent) : 0.f;
            float w10 = (va && k1ok) ? influence(spa, k1x, k1y, k1z, inv_extent) : 0.f;
            float w01 = (vb && k0ok) ? influence(spb, k0x, k0y, k0z, inv_extent) : 0.f;
            float w11 = (vb && k1ok) ? influence(spb, k1x, k1y, k1z, inv_extent) : 0.f;
            const float wv[4] = { w00, w10, w01, w11 };
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t hi = __float_as_uint(wv[u]) & TF32_MASK;
                ahi[s][u] = hi;
                alo[s][u] = __float_as_uint(wv[u] - __uint_as_float(hi));
            }
        }
#pragma unroll
        for (int s = 0; s < MMA_KS; s++) {
            if (h0 + 8 * s >= H) break;                    // warp-uniform
            float fa[NT], fb[NT];
            const float* ra = xl + (size_t)(ja[s] * (unsigned)ldx);
            const float* rb = xl + (size_t)(jb[s] * (unsigned)ldx);
            if (NT % 4 == 0) {
#pragma unroll
                for (int v = 0; v < NT / 4; v++) {
                    const float4 A = __ldg(reinterpret_cast<const float4*>(ra) + v);
                    const float4 B = __ldg(reinterpret_cast<const float4*>(rb) + v);
                    fa[4 * v] = A.x; fa[4 * v + 1] = A.y; fa[4 * v + 2] = A.z; fa[4 * v + 3] = A.w;
                    fb[4 * v] = B.x; fb[4 * v + 1] = B.y; fb[4 * v + 2] = B.z; fb[4 * v + 3] = B.w;
                }
            } else {
#pragma unroll
                for (int v = 0; v < NT; v++) { fa[v] = __ldg(ra + v); fb[v] = __ldg(rb + v); }
            }
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const uint32_t b0h = __float_as_uint(fa[nt]) & TF32_MASK, b1h = __float_as_uint(fb[nt]) & TF32_MASK;
                const uint32_t b0l = __float_as_uint(fa[nt] - __uint_as_float(b0h));
                const uint32_t b1l = __float_as_uint(fb[nt] - __uint_as_float(b1h));
                mma_tf32(acc[nt], alo[s], b0h, b1h);
                mma_tf32(acc[nt], ahi[s], b0l, b1l);
                mma_tf32(acc[nt], ahi[s], b0h, b1h);
            }
        }
    }
    // D fragment: acc[nt][0|1] -> kernel point g, channels c0 + (2t|2t+1)*NT + nt ; acc[nt][2|3] -> kernel point g+8
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int kp = g + 8 * r;
        if (kp >= K) continue;
        const size_t e0 = (size_t)n * ldk + (size_t)kp * cin + c0 + 2 * t * NT;
        if (SPLIT) {
            __align__(16) __nv_bfloat16 hi[2 * NT], lo[2 * NT];
#pragma unroll
            for (int half = 0; half < 2; half++)
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    const float v = acc[nt][2 * r + half];
                    const __nv_bfloat16 h = __float2bfloat16_rn(v);
                    hi[half * NT + nt] = h;
                    lo[half * NT + nt] = __float2bfloat16_rn(v - __bfloat162float(h));
                }
            if (NT >= 4) {
#pragma unroll
                for (int v = 0; v < (2 * NT) / 8; v++) {
                    *reinterpret_cast<uint4*>(wf_hi + e0 + 8 * v) = *reinterpret_cast<const uint4*>(hi + 8 * v);
                    *reinterpret_cast<uint4*>(wf_lo + e0 + 8 * v) = *reinterpret_cast<const uint4*>(lo + 8 * v);
                }
            } else if (NT == 2) {
                *reinterpret_cast<uint2*>(wf_hi + e0) = *reinterpret_cast<const uint2*>(hi);
                *reinterpret_cast<uint2*>(wf_lo + e0) = *reinterpret_cast<const uint2*>(lo);
            } else {
                *reinterpret_cast<uint32_t*>(wf_hi + e0) = *reinterpret_cast<const uint32_t*>(hi);
                *reinterpret_cast<uint32_t*>(wf_lo + e0) = *reinterpret_cast<const uint32_t*>(lo);
            }
        } else {
#pragma unroll
            for (int half = 0; half < 2; half++)
#pragma unroll
                for (int nt = 0; nt < NT; nt++) wf[e0 + half * NT + nt] = acc[nt][2 * r + half];
        }
    }
    if (blockIdx.y == 0) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
        if (lane == 0) inv_cnt[n] = 1.0f / (float)(cnt > 1 ? cnt : 1);
    }
}

// ---------------------------------------------------------------------------------------------
// 64-channel-slab variant of the tensor-core aggregation (cin % 64 == 0: every production layer).
// The neighbourhood of a point is first STAGED IN SHARED MEMORY with cp.async -- all (<= 40) feature
// rows of the slab and the neighbour coordinates are requested back to back, so the warp pays one
// memory round trip per point instead of one per k-step -- then the influence weights are computed
// just in time per k-step (no weight fragments kept live) and fed to mma.sync with the feature
// fragments read from shared memory (row stride 72 floats: conflict-free LDS.128 for the 4g / 32+4g
// channel mapping).
constexpr int A64_WARPS = 4;
constexpr int A64_ROWS = 24;
constexpr int A64_STRIDE = 72;
constexpr int A64_SMEM = A64_WARPS * (A64_ROWS * A64_STRIDE + A64_ROWS * 4) * 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, bool valid)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem), "r"(sz) : "memory");
}

template <typename IdxT, bool SPLIT>
__global__ void __launch_bounds__(A64_WARPS * 32) k_kpconv_aggregate_mma64(
    const float* __restrict__ q_pts, int nq, const float* __restrict__ s_pts, int ns, const IdxT* __restrict__ idx, int H,
    int idx_stride, const float* __restrict__ x, int cin, int ldx, const uint8_t* __restrict__ rowflag,
    const float* __restrict__ kpts, int K, float inv_extent, float* __restrict__ wf, __nv_bfloat16* __restrict__ wf_hi,
    __nv_bfloat16* __restrict__ wf_lo, int ldk, float* __restrict__ inv_cnt)
{
    extern __shared__ __align__(16) float smem_f[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = blockIdx.x * A64_WARPS + w;
    if (n >= nq) return;
    float* s_feat = smem_f + (size_t)w * (A64_ROWS * A64_STRIDE + A64_ROWS * 4);
    float* s_xyz = s_feat + A64_ROWS * A64_STRIDE;          // [row][x,y,z,valid]
    const int g = lane >> 2, t = lane & 3;
    const int c0 = blockIdx.y * 64;

    const float qx = q_pts[3 * (size_t)n], qy = q_pts[3 * (size_t)n + 1], qz = q_pts[3 * (size_t)n + 2];
    const bool k1ok = g + 8 < K, k0ok = g < K;
    const int ka = k0ok ? g : 0, kb = k1ok ? g + 8 : 0;
    const float k0x = kpts[3 * ka] + qx, k0y = kpts[3 * ka + 1] + qy, k0z = kpts[3 * ka + 2] + qz;
    const float k1x = kpts[3 * kb] + qx, k1y = kpts[3 * kb + 1] + qy, k1z = kpts[3 * kb + 2] + qz;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    int cnt = 0;
    const IdxT* row = idx + (size_t)n * idx_stride;
    const float* xl = x + c0 + (lane & 15) * 4;             // 16-byte chunk of the slab copied by this lane

    for (int h0 = 0; h0 < H; h0 += A64_ROWS) {
        // ---- stage: indices (coalesced), coordinates, flags, feature rows ----
        int ja = ns, jb = ns;
        if (lane < A64_ROWS && h0 + lane < H) { long long v = (long long)row[h0 + lane]; ja = (v >= 0 && v < ns) ? (int)v : ns; }
        if (lane < A64_ROWS - 32 && h0 + 32 + lane < H) { long long v = (long long)row[h0 + 32 + lane]; jb = (v >= 0 && v < ns) ? (int)v : ns; }
        const bool va = ja < ns, vb = jb < ns;
        if (lane < A64_ROWS) {
            float* d = s_xyz + lane * 4;
            const float* sp = s_pts + 3 * (size_t)(va ? ja : 0);
            cp_async4(d, sp, va); cp_async4(d + 1, sp + 1, va); cp_async4(d + 2, sp + 2, va);
            d[3] = va ? 1.f : 0.f;
            if (lane < A64_ROWS - 32) {
                float* d2 = s_xyz + (32 + lane) * 4;
                const float* sp2 = s_pts + 3 * (size_t)(vb ? jb : 0);
                cp_async4(d2, sp2, vb); cp_async4(d2 + 1, sp2 + 1, vb); cp_async4(d2 + 2, sp2 + 2, vb);
                d2[3] = vb ? 1.f : 0.f;
            }
        }
#pragma unroll 4
        for (int r = 0; r < A64_ROWS; r += 2) {
            const int rr = r + (lane >> 4);
            const int src_lane = rr & 31;
            const int j_lo = __shfl_sync(0xffffffffu, ja, src_lane), j_hi = __shfl_sync(0xffffffffu, jb, src_lane);
            const int j = rr < 32 ? j_lo : j_hi;
            const bool v = j < ns;
            cp_async16(s_feat + rr * A64_STRIDE + (lane & 15) * 4, xl + (size_t)((unsigned)(v ? j : 0) * (unsigned)ldx), v);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (blockIdx.y == 0) {
            const bool fa = va && rowflag[ja] != 0, fb = vb && rowflag[jb] != 0;
            cnt += __popc(__ballot_sync(0xffffffffu, fa)) + __popc(__ballot_sync(0xffffffffu, fb));
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();

        // ---- compute: one k-step = 8 neighbours ----
#pragma unroll
        for (int s = 0; s < A64_ROWS / 8; s++) {
            if (h0 + 8 * s >= H) break;                    // warp-uniform
            const float4 pa = *reinterpret_cast<const float4*>(s_xyz + (8 * s + t) * 4);
            const float4 pb = *reinterpret_cast<const float4*>(s_xyz + (8 * s + t + 4) * 4);
            float wv[4];
            {
                const float sa[3] = { pa.x, pa.y, pa.z }, sb[3] = { pb.x, pb.y, pb.z };
                wv[0] = (pa.w != 0.f && k0ok) ? influence(sa, k0x, k0y, k0z, inv_extent) : 0.f;
                wv[1] = (pa.w != 0.f && k1ok) ? influence(sa, k1x, k1y, k1z, inv_extent) : 0.f;
                wv[2] = (pb.w != 0.f && k0ok) ? influence(sb, k0x, k0y, k0z, inv_extent) : 0.f;
                wv[3] = (pb.w != 0.f && k1ok) ? influence(sb, k1x, k1y, k1z, inv_extent) : 0.f;
            }
            uint32_t ahi[4], alo[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                ahi[u] = __float_as_uint(wv[u]) & TF32_MASK;
                alo[u] = __float_as_uint(wv[u] - __uint_as_float(ahi[u]));
            }
            const float* ra = s_feat + (8 * s + t) * A64_STRIDE + 4 * g;
            const float* rb = s_feat + (8 * s + t + 4) * A64_STRIDE + 4 * g;
            const float4 A0 = *reinterpret_cast<const float4*>(ra), A1 = *reinterpret_cast<const float4*>(ra + 32);
            const float4 B0 = *reinterpret_cast<const float4*>(rb), B1 = *reinterpret_cast<const float4*>(rb + 32);
            const float fa[8] = { A0.x, A0.y, A0.z, A0.w, A1.x, A1.y, A1.z, A1.w };
            const float fb[8] = { B0.x, B0.y, B0.z, B0.w, B1.x, B1.y, B1.z, B1.w };
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                const uint32_t b0h = __float_as_uint(fa[nt]) & TF32_MASK, b1h = __float_as_uint(fb[nt]) & TF32_MASK;
                const uint32_t b0l = __float_as_uint(fa[nt] - __uint_as_float(b0h));
                const uint32_t b1l = __float_as_uint(fb[nt] - __uint_as_float(b1h));
                mma_tf32(acc[nt], alo, b0h, b1h);
                mma_tf32(acc[nt], ahi, b0l, b1l);
                mma_tf32(acc[nt], ahi, b0h, b1h);
            }
        }
        __syncwarp();
    }
    // tile nt < 4: column j <-> channel 4j + nt ; nt >= 4: channel 32 + 4j + (nt-4).  Lane holds columns 2t, 2t+1:
    // channels [8t, 8t+8) and [32+8t, 32+8t+8) for kernel points g (acc[.][0|1]) and g+8 (acc[.][2|3]).
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int kp = g + 8 * r;
        if (kp >= K) continue;
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {                     // hf = 0: channels 8t.. ; 1: 32+8t..
            float v[8];
#pragma unroll
            for (int c = 0; c < 8; c++) v[c] = acc[hf * 4 + (c & 3)][2 * r + (c >> 2)];
            const size_t e0 = (size_t)n * ldk + (size_t)kp * cin + c0 + hf * 32 + 8 * t;
            if (SPLIT) {
                __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    hi[c] = __float2bfloat16_rn(v[c]);
                    lo[c] = __float2bfloat16_rn(v[c] - __bfloat162float(hi[c]));
                }
                *reinterpret_cast<uint4*>(wf_hi + e0) = *reinterpret_cast<const uint4*>(hi);
                *reinterpret_cast<uint4*>(wf_lo + e0) = *reinterpret_cast<const uint4*>(lo);
            } else {
                *reinterpret_cast<float4*>(wf + e0) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(wf + e0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
    }
    if (blockIdx.y == 0 && lane == 0) inv_cnt[n] = 1.0f / (float)(cnt > 1 ? cnt : 1);
}

// ---------------------------------------------------------------------------------------------
// Few input channels (cin <= 4: the first layer has cin = 1): one THREAD per (query point, kernel point),
// 16 threads per point.  Each thread walks the neighbour list once and keeps its kernel point's cin
// accumulators; neighbour index / coordinates / features are broadcast loads inside the 16-thread group.
template <typename IdxT, int CIN, bool SPLIT>
__global__ void __launch_bounds__(256) k_kpconv_aggregate_small(
    const float* __restrict__ q_pts, int nq, const float* __restrict__ s_pts, int ns, const IdxT* __restrict__ idx, int H,
    int idx_stride, const float* __restrict__ x, int ldx, const uint8_t* __restrict__ rowflag, const float* __restrict__ kpts, int K,
    float inv_extent, float* __restrict__ wf, __nv_bfloat16* __restrict__ wf_hi, __nv_bfloat16* __restrict__ wf_lo, int ldk,
    float* __restrict__ inv_cnt)
{
    const int k = threadIdx.x & 15;
    const int n = blockIdx.x * 16 + (threadIdx.x >> 4);
    if (n >= nq) return;
    const bool kok = k < K;
    const int kk = kok ? k : 0;
    const float kx = kpts[3 * kk] + q_pts[3 * (size_t)n], ky = kpts[3 * kk + 1] + q_pts[3 * (size_t)n + 1],
                kz = kpts[3 * kk + 2] + q_pts[3 * (size_t)n + 2];
    float acc[CIN];
#pragma unroll
    for (int c = 0; c < CIN; c++) acc[c] = 0.f;
    int cnt = 0;
    const IdxT* row = idx + (size_t)n * idx_stride;
#pragma unroll 4
    for (int h = 0; h < H; h++) {
        const long long j = (long long)row[h];
        if (j < 0 || j >= ns) continue;                      // uniform inside the 16-thread group
        const float w = influence(s_pts + 3 * (size_t)j, kx, ky, kz, inv_extent);
        const float* xr = x + (size_t)j * ldx;
#pragma unroll
        for (int c = 0; c < CIN; c++) acc[c] = fmaf(w, __ldg(xr + c), acc[c]);
        cnt += rowflag[j];
    }
    if (kok) {
#pragma unroll
        for (int c = 0; c < CIN; c++) {
            const size_t e = (size_t)n * ldk + (size_t)k * CIN + c;
            if (SPLIT) {
                const __nv_bfloat16 h = __float2bfloat16_rn(acc[c]);
                wf_hi[e] = h;
                wf_lo[e] = __float2bfloat16_rn(acc[c] - __bfloat162float(h));
            } else {
                wf[e] = acc[c];
            }
        }
    }
    if (k == 0) inv_cnt[n] = 1.0f / (float)(cnt > 1 ? cnt : 1);
}

// ---------------------------------------------------------------------------------------------
// First layer (cin <= 4, e.g. the all-ones feature of the reference pipeline, cin = 1): the WHOLE KPConv in one kernel.
// 8 threads per query point (four points per warp, 32 per block).  (1) the point's neighbourhood is staged once in shared
// memory by its 8 threads: one float4 (p - q, first feature) per neighbour (+ one with the other features when cin > 1);
// (2) thread t accumulates kernel points t and t + 8 over the neighbours (one broadcast LDS.128 per neighbour and thread);
// (3) the [K*cin] x cout weight contraction: the aggregates travel by shuffle inside the 8-thread group, thread t produces
// the NJ = cout/8 consecutive output channels t*NJ.. from weights held in shared memory as [K*cin][8][NJ]; (4) x 1/neighbour
// count, row-contiguous store.  Replaces aggregate_small + the fp32 CUDA-core contraction and their [Nq, K*cin] round trip.
constexpr int SF_HMAX = 64;
constexpr int SF_PTS = 32;                        // points per block

template <typename IdxT, int CIN, int NJ>
__global__ void __launch_bounds__(256) k_kpconv_small_fused(
    const float* __restrict__ q_pts, int nq, const float* __restrict__ s_pts, int ns, const IdxT* __restrict__ idx, int H,
    int idx_stride, const float* __restrict__ x, int ldx, const uint8_t* __restrict__ rowflag, const float* __restrict__ kpts, int K,
    float inv_extent, const float* __restrict__ weights, float* __restrict__ out)
{
    extern __shared__ __align__(16) float smem_f[];
    constexpr int REC = CIN > 1 ? 2 : 1;                                 // float4 records per neighbour
    float4* s_nb = reinterpret_cast<float4*>(smem_f);                    // [SF_PTS][SF_HMAX][REC]
    float* s_w = smem_f + SF_PTS * SF_HMAX * REC * 4;                    // [K*CIN][NJ/4][8][4]  (see the staging loop)
    constexpr int COUT = 8 * NJ;
    const int t = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const int n = blockIdx.x * SF_PTS + pl;
    for (int e = threadIdx.x; e < K * CIN * COUT; e += 256) {
        const int kc = e / COUT, c = e - kc * COUT;                      // weights [K][CIN][COUT] row-major; channel c = t*NJ + j
        const int tt = c / NJ, j = c % NJ;
        // [kc][j / 4][t][4] when NJ % 4 == 0: the 8 threads of a point read 8 adjacent float4 (bank-conflict free); else [kc][t][NJ]
        s_w[NJ % 4 == 0 ? ((kc * (NJ / 4) + j / 4) * 8 + tt) * 4 + (j & 3) : (kc * 8 + tt) * NJ + j] = weights[e];
    }
    const bool live = n < nq;
    const int nn = live ? n : nq - 1;
    const float qx = q_pts[3 * (size_t)nn], qy = q_pts[3 * (size_t)nn + 1], qz = q_pts[3 * (size_t)nn + 2];
    float4* nb = s_nb + pl * (SF_HMAX * REC);
    int cnt = 0;
    const IdxT* row = idx + (size_t)nn * idx_stride;
    for (int h = t; h < H; h += 8) {
        const long long j = (long long)row[h];
        float4 a = make_float4(1e15f, 1e15f, 1e15f, 0.f), f = make_float4(0.f, 0.f, 0.f, 0.f);    // shadow: influence 0, zero features
        if (j >= 0 && j < ns) {
            const float* sp = s_pts + 3 * (size_t)j;
            const float* xr = x + (size_t)j * ldx;
            a = make_float4(sp[0] - qx, sp[1] - qy, sp[2] - qz, xr[0]);
            if (CIN > 1) f.x = xr[1];
            if (CIN > 2) f.y = xr[2];
            if (CIN > 3) f.z = xr[3];
            cnt += rowflag[j];
        }
        nb[REC * h] = a;
        if (CIN > 1) nb[REC * h + 1] = f;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o, 8);
    __syncthreads();                                                     // weights + every point's neighbourhood staged

    // (2) aggregates of kernel points t and t + 8 (a missing kernel point sits far away: influence exactly 0)
    const bool ok0 = t < K, ok1 = t + 8 < K;
    const float k0x = ok0 ? kpts[3 * t] : 1e15f, k0y = ok0 ? kpts[3 * t + 1] : 1e15f, k0z = ok0 ? kpts[3 * t + 2] : 1e15f;
    const float k1x = ok1 ? kpts[3 * (t + 8)] : 1e15f, k1y = ok1 ? kpts[3 * (t + 8) + 1] : 1e15f, k1z = ok1 ? kpts[3 * (t + 8) + 2] : 1e15f;
    float acc[2][CIN];
#pragma unroll
    for (int c = 0; c < CIN; c++) acc[0][c] = acc[1][c] = 0.f;
#pragma unroll 4
    for (int h = 0; h < H; h++) {
        const float4 a = nb[REC * h];
        float f[4] = { a.w, 0.f, 0.f, 0.f };
        if (CIN > 1) { const float4 g = nb[REC * h + 1]; f[1] = g.x; f[2] = g.y; f[3] = g.z; }
        float dx = a.x - k0x, dy = a.y - k0y, dz = a.z - k0z;
        float d2 = fmaxf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)), 1e-30f);
        const float w0 = fmaxf(0.f, fmaf(-d2 * rsqrtf(d2), inv_extent, 1.f));
        dx = a.x - k1x; dy = a.y - k1y; dz = a.z - k1z;
        d2 = fmaxf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)), 1e-30f);
        const float w1 = fmaxf(0.f, fmaf(-d2 * rsqrtf(d2), inv_extent, 1.f));
#pragma unroll
        for (int c = 0; c < CIN; c++) { acc[0][c] = fmaf(w0, f[c], acc[0][c]); acc[1][c] = fmaf(w1, f[c], acc[1][c]); }
    }
    // (3) contraction: out[c] = sum_{kp,ci} wf[kp][ci] * W[kp][ci][c] for this thread's NJ channels
    float o[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++) o[j] = 0.f;
    for (int kp = 0; kp < K; kp++) {
#pragma unroll
        for (int ci = 0; ci < CIN; ci++) {
            const float lo = __shfl_sync(0xffffffffu, acc[0][ci], kp & 7, 8), hi = __shfl_sync(0xffffffffu, acc[1][ci], kp & 7, 8);
            const float v = kp < 8 ? lo : hi;
            const float* wr = s_w + ((kp * CIN + ci) * 8 + t) * NJ;
            if (NJ % 4 == 0) {
                const float4* w4p = reinterpret_cast<const float4*>(s_w) + (kp * CIN + ci) * (NJ / 4) * 8 + t;
#pragma unroll
                for (int j = 0; j + 3 < NJ; j += 4) {
                    const float4 w4 = w4p[(j / 4) * 8];
                    o[j] = fmaf(v, w4.x, o[j]); o[j + 1] = fmaf(v, w4.y, o[j + 1]);
                    o[j + 2] = fmaf(v, w4.z, o[j + 2]); o[j + 3] = fmaf(v, w4.w, o[j + 3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NJ; j++) o[j] = fmaf(v, wr[j], o[j]);
            }
        }
    }
    if (live) {
        const float sc = 1.0f / (float)(cnt > 1 ? cnt : 1);                // models/blocks.py:369-372
        float* dst = out + (size_t)n * COUT + t * NJ;
        if (NJ % 4 == 0) {
#pragma unroll
            for (int j = 0; j + 3 < NJ; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(o[j] * sc, o[j + 1] * sc, o[j + 2] * sc, o[j + 3] * sc);
        } else {
#pragma unroll
            for (int j = 0; j < NJ; j++) dst[j] = o[j] * sc;
        }
    }
}

template <typename IdxT>
static bool launch_small_fused(const float* q_pts, int nq, const float* s_pts, int ns, const IdxT* idx, int H, int idx_stride, const float* x,
                               int cin, const uint8_t* rowflag, const float* kpts, int K, float inv_extent, const float* weights, int cout,
                               float* out, cudaStream_t st)
{
    if (cin > 4 || H > SF_HMAX || (cout != 16 && cout != 32 && cout != 64 && cout != 128 && cout != 256)) return false;
    const size_t smem = (size_t)SF_PTS * SF_HMAX * (cin > 1 ? 2 : 1) * 16 + (size_t)K * cin * cout * sizeof(float);
    if (smem > 160 * 1024) return false;
    const unsigned grid = (unsigned)cdiv64(nq, SF_PTS);
#define PCRCG_SF(C_, NJ_)                                                                                                              \
    do {                                                                                                                               \
        cudaFuncSetAttribute(k_kpconv_small_fused<IdxT, C_, NJ_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
        k_kpconv_small_fused<IdxT, C_, NJ_><<<grid, 256, smem, st>>>(q_pts, nq, s_pts, ns, idx, H, idx_stride, x, cin, rowflag, kpts, K, \
                                                                     inv_extent, weights, out);                                       \
    } while (0)
#define PCRCG_SF_C(C_)                                                                                            \
    do {                                                                                                          \
        if (cout == 16) PCRCG_SF(C_, 2); else if (cout == 32) PCRCG_SF(C_, 4); else if (cout == 64) PCRCG_SF(C_, 8); \
        else if (cout == 128) PCRCG_SF(C_, 16); else PCRCG_SF(C_, 32);                                             \
    } while (0)
    if (cin == 1) PCRCG_SF_C(1);
    else if (cin == 2) PCRCG_SF_C(2);
    else if (cin == 3) PCRCG_SF_C(3);
    else PCRCG_SF_C(4);
#undef PCRCG_SF_C
#undef PCRCG_SF
    return true;
}

// ---------------------------------------------------------------------------------------------
// bf16x3 variant of the 64-channel-slab tensor-core aggregation, used when the producer of the features
// already emitted them as bf16 (hi, lo) planes (instance norm epilogue, dense.cu).  Compared with the
// 3xTF32 kernel above: mma.sync m16n8k16 (16 neighbours per step: half the MMAs), B fragments come
// straight out of shared memory with ldmatrix.trans (no per-use hi/lo splitting), and the result tile is
// transposed through shared memory with stmatrix so the bf16 planes are written in whole 128-byte rows.
//   A (influence weights, split hi/lo in registers) x B (feature planes hi / lo):  hi*hi + hi*lo + lo*hi
constexpr int AB_WARPS = 4;
constexpr int AB_ROWS = 32;                    // neighbours staged per chunk = 2 k-steps of 16
constexpr int AB_WARP_BYTES = 2 * AB_ROWS * AB_PITCH + AB_ROWS * 16;    // hi plane, lo plane, coordinates
constexpr int AB_SMEM = AB_WARPS * AB_WARP_BYTES;

template <typename IdxT>
__global__ void __launch_bounds__(AB_WARPS * 32) k_kpconv_aggregate_bf16(
    const float* __restrict__ q_pts, int nq, const float* __restrict__ s_pts, int ns, const IdxT* __restrict__ idx, int H,
    int idx_stride, const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo, int cin, int ldxs,
    const uint8_t* __restrict__ rowflag, const float* __restrict__ kpts, int K, float inv_extent,
    __nv_bfloat16* __restrict__ wf_hi, __nv_bfloat16* __restrict__ wf_lo, int ldk, float* __restrict__ inv_cnt)
{
    extern __shared__ __align__(16) uint8_t smem_b[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = blockIdx.x * AB_WARPS + w;
    if (n >= nq) return;
    uint8_t* s_hi = smem_b + (size_t)w * AB_WARP_BYTES;
    uint8_t* s_lo = s_hi + AB_ROWS * AB_PITCH;
    float* s_xyz = reinterpret_cast<float*>(s_lo + AB_ROWS * AB_PITCH);      // [row][x', y', z', |p'|^2 or 1e30 for a shadow]  (p' = p - q)
    const uint32_t a_hi = (uint32_t)__cvta_generic_to_shared(s_hi), a_lo = (uint32_t)__cvta_generic_to_shared(s_lo);
    const int g = lane >> 2, t = lane & 3;
    const int c0 = blockIdx.y * 64;

    const float qx = q_pts[3 * (size_t)n], qy = q_pts[3 * (size_t)n + 1], qz = q_pts[3 * (size_t)n + 2];
    // |p' - k|^2 = |p'|^2 + |k|^2 - 2 p'.k : the kernel-point terms are per-lane constants (kernel points g and g+8)
    const bool k1ok = g + 8 < K, k0ok = g < K;
    const int ka = k0ok ? g : 0, kb = k1ok ? g + 8 : 0;
    const float k0x = -2.f * kpts[3 * ka], k0y = -2.f * kpts[3 * ka + 1], k0z = -2.f * kpts[3 * ka + 2];
    const float k1x = -2.f * kpts[3 * kb], k1y = -2.f * kpts[3 * kb + 1], k1z = -2.f * kpts[3 * kb + 2];
    // a missing kernel point (k >= K) or a shadow neighbour gets a huge squared distance -> influence exactly 0
    const float k0n = k0ok ? 0.25f * (k0x * k0x + k0y * k0y + k0z * k0z) : 1e30f;
    const float k1n = k1ok ? 0.25f * (k1x * k1x + k1y * k1y + k1z * k1z) : 1e30f;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    int cnt = 0;
    const IdxT* row = idx + (size_t)n * idx_stride;
    // staging: lanes 0..7 copy the 8 x 16-byte chunks of a hi row, 8..15 of the lo row; lanes 16..31 the next neighbour
    const int chunk = lane & 7, plane = (lane >> 3) & 1, rsel = lane >> 4;
    const __nv_bfloat16* xp = (plane ? x_lo : x_hi) + c0 + chunk * 8;
    const uint32_t dst0 = (plane ? a_lo : a_hi) + (uint32_t)(chunk * 16 + rsel * AB_PITCH);
    // ldmatrix lane addressing: matrix m = lane>>3 -> (k half = m&1, n-tile offset = m>>1), row = lane&7
    const uint32_t lm_off = (uint32_t)(((lane >> 3) & 1) * 8 + (lane & 7)) * AB_PITCH + (uint32_t)(lane >> 4) * 16;

    for (int h0 = 0; h0 < H; h0 += AB_ROWS) {
        int ja = ns;
        if (h0 + lane < H) { long long v = (long long)row[h0 + lane]; ja = (v >= 0 && v < ns) ? (int)v : ns; }
        const bool va = ja < ns;
        // Neighbour lists are padded with shadows: a 16-neighbour k-step made of shadows only contributes exactly 0
        // (zero-filled rows x zero weights) and is neither staged nor multiplied.  vmask is warp-uniform.
        const uint32_t vmask = __ballot_sync(0xffffffffu, va);
        if (vmask == 0u) continue;
        const bool second = (vmask >> 16) != 0u;
        // feature planes: 16 unrolled cp.async of 16 bytes per lane (2 neighbour rows per step)
#pragma unroll
        for (int r = 0; r < 16; r += 2) {
            const int j = __shfl_sync(0xffffffffu, ja, r + rsel);
            const bool v = j < ns;
            const void* src = xp + (size_t)((unsigned)(v ? j : 0) * (unsigned)ldxs);
            const int sz = v ? 16 : 0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (uint32_t)(r * AB_PITCH)), "l"(src), "r"(sz) : "memory");
        }
        if (second) {
#pragma unroll
            for (int r = 16; r < AB_ROWS; r += 2) {
                const int j = __shfl_sync(0xffffffffu, ja, r + rsel);
                const bool v = j < ns;
                const void* src = xp + (size_t)((unsigned)(v ? j : 0) * (unsigned)ldxs);
                const int sz = v ? 16 : 0;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (uint32_t)(r * AB_PITCH)), "l"(src), "r"(sz) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // coordinates relative to the query, squared norm (or -1 = shadow) -- one neighbour per lane
        {
            float px = 0.f, py = 0.f, pz = 0.f, pn = 1e30f;
            if (va) {
                const float* sp = s_pts + 3 * (size_t)ja;
                px = sp[0] - qx; py = sp[1] - qy; pz = sp[2] - qz;
                pn = fmaf(px, px, fmaf(py, py, pz * pz));
            }
            *reinterpret_cast<float4*>(s_xyz + lane * 4) = make_float4(px, py, pz, pn);
        }
        if (blockIdx.y == 0) cnt += __popc(__ballot_sync(0xffffffffu, va && rowflag[ja] != 0));
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();

#pragma unroll
        for (int s = 0; s < AB_ROWS / 16; s++) {
            if (((vmask >> (16 * s)) & 0xffffu) == 0u) continue;          // warp-uniform: nothing but shadows in this k-step
            // A fragment: kernel points (g, g+8) x neighbours (2t, 2t+1 | 2t+8, 2t+9) of this 16-neighbour step
            float wv[2][4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int rn = 16 * s + 2 * t + (e & 1) + (e >> 1) * 8;
                const float4 p = *reinterpret_cast<const float4*>(s_xyz + rn * 4);
                // the expanded squared distance can come out slightly negative next to a kernel point: |.| (a free operand
                // modifier) instead of a max keeps the square root defined; the error is the same few ulps of |p'|^2
                const float d0 = fabsf(fmaf(p.x, k0x, fmaf(p.y, k0y, fmaf(p.z, k0z, p.w + k0n))));
                const float d1 = fabsf(fmaf(p.x, k1x, fmaf(p.y, k1y, fmaf(p.z, k1z, p.w + k1n))));
                wv[0][e] = fmaxf(0.f, fmaf(-sqrt_approx(d0), inv_extent, 1.f));
                wv[1][e] = fmaxf(0.f, fmaf(-sqrt_approx(d1), inv_extent, 1.f));
            }
            uint32_t ahi[4], alo[4];
            ahi[0] = pack_split(wv[0][0], wv[0][1], alo[0]);     // row g   , k 2t..2t+1
            ahi[1] = pack_split(wv[1][0], wv[1][1], alo[1]);     // row g+8 , k 2t..2t+1
            ahi[2] = pack_split(wv[0][2], wv[0][3], alo[2]);     // row g   , k 2t+8..2t+9
            ahi[3] = pack_split(wv[1][2], wv[1][3], alo[3]);     // row g+8 , k 2t+8..2t+9
            const uint32_t sbase = (uint32_t)(16 * s) * AB_PITCH + lm_off;
#pragma unroll
            for (int np = 0; np < 4; np++) {                      // pairs of n-tiles (2 x 8 channels = 32 bytes)
                uint32_t bh[4], bl[4];
                ldmatrix_x4_trans(bh, a_hi + sbase + np * 32);
                ldmatrix_x4_trans(bl, a_lo + sbase + np * 32);
                mma_bf16(acc[2 * np], alo, bh[0], bh[1]);
                mma_bf16(acc[2 * np + 1], alo, bh[2], bh[3]);
                mma_bf16(acc[2 * np], ahi, bl[0], bl[1]);
                mma_bf16(acc[2 * np + 1], ahi, bl[2], bl[3]);
                mma_bf16(acc[2 * np], ahi, bh[0], bh[1]);
                mma_bf16(acc[2 * np + 1], ahi, bh[2], bh[3]);
            }
        }
        __syncwarp();
    }
    // D tile [16 kp x 64 ch] -> bf16 hi / lo -> shared (stmatrix, row pitch AB_PITCH) -> global rows of 128 bytes.
    // stmatrix.x4: matrix m = lane>>3 supplies row addresses; we store (kp 0-7 | 8-15) x (n-tile 2np, 2np+1).
    {
        const uint32_t st_off = (uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8) * AB_PITCH + (uint32_t)(lane >> 4) * 16;
#pragma unroll
        for (int np = 0; np < 4; np++) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {        // q: 0 = (rows 0-7, tile 2np), 1 = (rows 8-15, tile 2np), 2 = (rows 0-7, tile 2np+1), 3 = (rows 8-15, 2np+1)
                const int tile = 2 * np + (q >> 1), rh = q & 1;
                h[q] = pack_split(acc[tile][2 * rh], acc[tile][2 * rh + 1], l[q]);
            }
            stmatrix_x4(a_hi + st_off + np * 32, h[0], h[1], h[2], h[3]);
            stmatrix_x4(a_lo + st_off + np * 32, l[0], l[1], l[2], l[3]);
        }
        __syncwarp();
        // 16 rows x 128 bytes per plane: 8 lanes per row, 4 rows per instruction
        const int rl = lane >> 3, cl = lane & 7;
        const size_t ebase = (size_t)n * ldk + c0 + cl * 8;
#pragma unroll
        for (int r0 = 0; r0 < 16; r0 += 4) {
            const int kp = r0 + rl;
            if (kp < K) {
                const size_t e = ebase + (size_t)kp * cin;
                *reinterpret_cast<uint4*>(wf_hi + e) = *reinterpret_cast<const uint4*>(s_hi + kp * AB_PITCH + cl * 16);
                *reinterpret_cast<uint4*>(wf_lo + e) = *reinterpret_cast<const uint4*>(s_lo + kp * AB_PITCH + cl * 16);
            }
        }
    }
    if (blockIdx.y == 0 && lane == 0) inv_cnt[n] = 1.0f / (float)(cnt > 1 ? cnt : 1);
}

// ---------------------------------------------------------------------------------------------
// Persistent, software-pipelined form of the bf16x3 aggregation (H <= 64).  The one-point-per-warp kernel above spends a
// third of its issue cycles waiting on L2 (long scoreboard: neighbour indices -> feature rows -> coordinates are three
// dependent round trips per point, hidden only by the ~17 other resident warps).  Here a warp walks over many points and
// keeps one 16-neighbour k-step in flight while it multiplies the previous one:
//     iteration i:  cp.async of k-step i+1 (other buffer)  |  coordinates of k-step i+1 -> registers  |  wait k-step i
//                   MMAs of k-step i  |  coordinates -> shared  |  (last k-step of a point: transposed store, which uses
//                   the buffer just consumed as scratch)
// The neighbour indices (and row flags) of the NEXT point are requested when the current point starts.  Shared memory
// per warp is unchanged (2 buffers x 16 rows instead of 1 x 32), so residency stays at 5 CTAs / SM.
constexpr int ABP_SMEM = AB_WARPS * ABP_WARP_BYTES;

template <typename IdxT>
__global__ void __launch_bounds__(AB_WARPS * 32, 5) k_kpconv_aggregate_bf16p(
    const float* __restrict__ q_pts, int nq, const float* __restrict__ s_pts, int ns, const IdxT* __restrict__ idx, int H,
    int idx_stride, const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo, int cin, int ldxs,
    const uint8_t* __restrict__ rowflag, const float* __restrict__ kpts, int K, float inv_extent,
    __nv_bfloat16* __restrict__ wf_hi, __nv_bfloat16* __restrict__ wf_lo, int ldk, float* __restrict__ inv_cnt)
{
    extern __shared__ __align__(16) uint8_t smem_b[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int stride = gridDim.x * AB_WARPS;
    int n = blockIdx.x * AB_WARPS + w;
    if (n >= nq) return;
    uint8_t* s_buf = smem_b + (size_t)w * ABP_WARP_BYTES;                    // buffer b: hi rows at b*BUF, lo rows at b*BUF + 16*PITCH
    float* s_xyz = reinterpret_cast<float*>(s_buf + 2 * ABP_BUF_BYTES);       // [2][16][4]
    const uint32_t a_buf = (uint32_t)__cvta_generic_to_shared(s_buf);
    const int g = lane >> 2, t = lane & 3;
    const int c0 = blockIdx.y * 64;
    const bool k1ok = g + 8 < K, k0ok = g < K;
    const int ka = k0ok ? g : 0, kb = k1ok ? g + 8 : 0;
    const float k0x = -2.f * kpts[3 * ka], k0y = -2.f * kpts[3 * ka + 1], k0z = -2.f * kpts[3 * ka + 2];
    const float k1x = -2.f * kpts[3 * kb], k1y = -2.f * kpts[3 * kb + 1], k1z = -2.f * kpts[3 * kb + 2];
    const float k0n = k0ok ? 0.25f * (k0x * k0x + k0y * k0y + k0z * k0z) : 1e30f;
    const float k1n = k1ok ? 0.25f * (k1x * k1x + k1y * k1y + k1z * k1z) : 1e30f;
    const int chunk = lane & 7, plane = (lane >> 3) & 1, rsel = lane >> 4;
    const __nv_bfloat16* xp = (plane ? x_lo : x_hi) + c0 + chunk * 8;
    const uint32_t dst_off = (uint32_t)(plane * ABP_ROWS * AB_PITCH + chunk * 16 + rsel * AB_PITCH);
    const uint32_t lm_off = (uint32_t)(((lane >> 3) & 1) * 8 + (lane & 7)) * AB_PITCH + (uint32_t)(lane >> 4) * 16;

    // neighbour indices of a point: the RAW values are requested early (load_raw) and only clamped to the shadow convention
    // when the point is about to start (clamp_idx), so the load latency hides behind the previous point's k-steps
    auto load_raw = [&](int p, IdxT& r0, IdxT& r1) {
        const IdxT* row = idx + (size_t)min(p, nq - 1) * idx_stride;
        r0 = row[lane < H ? lane : 0];
        r1 = row[lane + 32 < H ? lane + 32 : 0];
    };
    auto clamp_idx = [&](int p, IdxT r0, IdxT r1, int& j0, int& j1) {
        const long long v0 = (long long)r0, v1 = (long long)r1;
        j0 = (p < nq && lane < H && v0 >= 0 && v0 < ns) ? (int)v0 : ns;
        j1 = (p < nq && lane + 32 < H && v1 >= 0 && v1 < ns) ? (int)v1 : ns;
    };
    // k-steps (of 16 neighbours) that hold at least one real neighbour, as a 4-bit mask
    auto step_mask = [&](int j0, int j1) -> uint32_t {
        const uint32_t m0 = __ballot_sync(0xffffffffu, j0 < ns), m1 = __ballot_sync(0xffffffffu, j1 < ns);
        return ((m0 & 0xffffu) ? 1u : 0u) | ((m0 >> 16) ? 2u : 0u) | ((m1 & 0xffffu) ? 4u : 0u) | ((m1 >> 16) ? 8u : 0u);
    };
    // stage k-step s of a point (neighbour registers j0 / j1) into buffer b: 8 cp.async per lane, 2 rows per instruction
    auto stage = [&](int j0, int j1, int s, int b) {
        const int jsrc = s < 2 ? j0 : j1;
        const uint32_t dst = a_buf + (uint32_t)(b * ABP_BUF_BYTES) + dst_off;
#pragma unroll
        for (int r = 0; r < ABP_ROWS; r += 2) {
            const int j = __shfl_sync(0xffffffffu, jsrc, ((s & 1) << 4) + r + rsel);
            const bool v = j < ns;
            const void* src = xp + (size_t)((unsigned)(v ? j : 0) * (unsigned)ldxs);
            const int sz = v ? 16 : 0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)(r * AB_PITCH)), "l"(src), "r"(sz) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // raw coordinates of the neighbour this lane (< 16) owns in k-step s
    auto load_xyz = [&](int j0, int j1, int s, float& x, float& y, float& z) -> bool {
        const int j = __shfl_sync(0xffffffffu, s < 2 ? j0 : j1, ((s & 1) << 4) + (lane & 15));
        const bool v = j < ns;
        if (v && lane < 16) { const float* sp = s_pts + 3 * (size_t)j; x = sp[0]; y = sp[1]; z = sp[2]; }
        return v;
    };

    // write the coordinate block of a staged k-step: p' = p - q and |p'|^2 (1e30 for a shadow: influence exactly 0)
    auto put_xyz = [&](int blk, bool v, float x, float y, float z, float ox, float oy, float oz) {
        if (lane < 16) {
            const float px = x - ox, py = y - oy, pz = z - oz;
            *reinterpret_cast<float4*>(s_xyz + blk * (ABP_ROWS * 4) + lane * 4) =
                v ? make_float4(px, py, pz, fmaf(px, px, fmaf(py, py, pz * pz))) : make_float4(0.f, 0.f, 0.f, 1e30f);
        }
    };
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

    // multiply the k-step held by buffer blk into acc
    auto compute = [&](int blk) {
        const float* xyz = s_xyz + blk * (ABP_ROWS * 4);
        float wv[2][4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int rn = 2 * t + (e & 1) + (e >> 1) * 8;
            const float4 p = *reinterpret_cast<const float4*>(xyz + rn * 4);
            const float d0 = fabsf(fmaf(p.x, k0x, fmaf(p.y, k0y, fmaf(p.z, k0z, p.w + k0n))));
            const float d1 = fabsf(fmaf(p.x, k1x, fmaf(p.y, k1y, fmaf(p.z, k1z, p.w + k1n))));
            wv[0][e] = fmaxf(0.f, fmaf(-sqrt_approx(d0), inv_extent, 1.f));
            wv[1][e] = fmaxf(0.f, fmaf(-sqrt_approx(d1), inv_extent, 1.f));
        }
        uint32_t ahi[4], alo[4];
        ahi[0] = pack_split(wv[0][0], wv[0][1], alo[0]);
        ahi[1] = pack_split(wv[1][0], wv[1][1], alo[1]);
        ahi[2] = pack_split(wv[0][2], wv[0][3], alo[2]);
        ahi[3] = pack_split(wv[1][2], wv[1][3], alo[3]);
        const uint32_t base_hi = a_buf + (uint32_t)(blk * ABP_BUF_BYTES) + lm_off, base_lo = base_hi + ABP_ROWS * AB_PITCH;
#pragma unroll
        for (int np = 0; np < 4; np++) {
            uint32_t bh[4], bl[4];
            ldmatrix_x4_trans(bh, base_hi + np * 32);
            ldmatrix_x4_trans(bl, base_lo + np * 32);
            mma_bf16(acc[2 * np], alo, bh[0], bh[1]);
            mma_bf16(acc[2 * np + 1], alo, bh[2], bh[3]);
            mma_bf16(acc[2 * np], ahi, bl[0], bl[1]);
            mma_bf16(acc[2 * np + 1], ahi, bl[2], bl[3]);
            mma_bf16(acc[2 * np], ahi, bh[0], bh[1]);
            mma_bf16(acc[2 * np + 1], ahi, bh[2], bh[3]);
        }
    };
    // D [16 kp x 64 ch] of point pt -> bf16 hi / lo planes, straight from the accumulator registers: inside the slab the
    // channels are stored in kperm64 order (common.cuh), which makes a thread's tiles 0-3 / 4-7 one 16-byte piece each and
    // the four threads of a kernel-point row 64 contiguous bytes -- no shared-memory transposition.  acc is cleared.
    auto store_point = [&](int pt) {
        const size_t ebase = (size_t)pt * ldk + c0 + 8 * t;
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            const int kp = g + 8 * hh;
            uint32_t h[8], l[8];
#pragma unroll
            for (int nt = 0; nt < 8; nt++) h[nt] = pack_split(acc[nt][2 * hh], acc[nt][2 * hh + 1], l[nt]);
            if (kp < K) {
                const size_t e = ebase + (size_t)kp * cin;
                *reinterpret_cast<uint4*>(wf_hi + e) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(wf_hi + e + 32) = make_uint4(h[4], h[5], h[6], h[7]);
                *reinterpret_cast<uint4*>(wf_lo + e) = make_uint4(l[0], l[1], l[2], l[3]);
                *reinterpret_cast<uint4*>(wf_lo + e + 32) = make_uint4(l[4], l[5], l[6], l[7]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    };

    // ---- prologue: first point, its first k-step into buffer 0 --------------------------------------------------
    int j0, j1;
    {
        IdxT r0, r1;
        load_raw(n, r0, r1);
        clamp_idx(n, r0, r1, j0, j1);
    }
    uint32_t smask = step_mask(j0, j1);
    int s = smask ? __ffs(smask) - 1 : -1;
    float qx = q_pts[3 * (size_t)n], qy = q_pts[3 * (size_t)n + 1], qz = q_pts[3 * (size_t)n + 2];
    int b = 0;
    if (s >= 0) {
        stage(j0, j1, s, 0);
        float x = 0.f, y = 0.f, z = 0.f;
        const bool v = load_xyz(j0, j1, s, x, y, z);
        put_xyz(0, v, x, y, z, qx, qy, qz);
    }

    IdxT nr0, nr1;                               // raw neighbour indices of the next point (requested one point earlier)
    load_raw(n + stride, nr0, nr1);
    while (n < nq) {
        // requested now: neighbour indices of the point AFTER the next one, query of the next point, row flags of this one
        const int nn = n + stride;
        IdxT fr0, fr1;
        load_raw(nn + stride, fr0, fr1);
        int nj0 = ns, nj1 = ns;
        const size_t qo = 3 * (size_t)min(nn, nq - 1);
        const float nqx = q_pts[qo], nqy = q_pts[qo + 1], nqz = q_pts[qo + 2];
        uint8_t f0 = 0, f1 = 0;                  // row flags of this point's neighbours (compared when the point ends)
        if (blockIdx.y == 0) { f0 = rowflag[j0 < ns ? j0 : 0]; f1 = rowflag[j1 < ns ? j1 : 0]; }
        uint32_t nmask = 0;
        int ns_first = -1;                       // first real k-step of the next point, once known

        if (s < 0) {
            // a point without any real neighbour: zero rows; nothing of the next point is in flight yet
            store_point(n);
            clamp_idx(nn, nr0, nr1, nj0, nj1);
            nmask = step_mask(nj0, nj1);
            ns_first = (nn < nq && nmask) ? __ffs(nmask) - 1 : -1;
            if (ns_first >= 0) {
                stage(nj0, nj1, ns_first, b);
                float x = 0.f, y = 0.f, z = 0.f;
                const bool v = load_xyz(nj0, nj1, ns_first, x, y, z);
                put_xyz(b, v, x, y, z, nqx, nqy, nqz);
            }
        } else {
            while (true) {
                // what follows (n, s): the next real k-step of this point, else the first one of the next point
                const uint32_t rest = smask & ~((2u << s) - 1u);
                const bool same = rest != 0u;
                int s2;
                if (same) s2 = __ffs(rest) - 1;
                else {
                    clamp_idx(nn, nr0, nr1, nj0, nj1);
                    nmask = step_mask(nj0, nj1);
                    ns_first = (nn < nq && nmask) ? __ffs(nmask) - 1 : -1;
                    s2 = ns_first;
                }
                float x = 0.f, y = 0.f, z = 0.f;
                bool v2 = false;
                if (s2 >= 0) {
                    stage(same ? j0 : nj0, same ? j1 : nj1, s2, b ^ 1);
                    v2 = load_xyz(same ? j0 : nj0, same ? j1 : nj1, s2, x, y, z);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                __syncwarp();
                compute(b);
                if (s2 >= 0) put_xyz(b ^ 1, v2, x, y, z, same ? qx : nqx, same ? qy : nqy, same ? qz : nqz);
                __syncwarp();
                if (!same) break;
                s = s2;
                b ^= 1;
            }
            store_point(n);
            b ^= 1;                              // the next point's first k-step is in flight in the other buffer
        }
        if (blockIdx.y == 0) {
            const int cnt = __popc(__ballot_sync(0xffffffffu, j0 < ns && f0 != 0)) + __popc(__ballot_sync(0xffffffffu, j1 < ns && f1 != 0));
            if (lane == 0) inv_cnt[n] = 1.0f / (float)(cnt > 1 ? cnt : 1);
        }
        n = nn;
        j0 = nj0; j1 = nj1;
        nr0 = fr0; nr1 = fr1;
        qx = nqx; qy = nqy; qz = nqz;
        smask = nmask;
        s = ns_first;
    }
}

// cin = 64 m + 1 (PCR-CG's colour input: 128 image-feature channels + 1, configs/test/indoor.yaml:34): the first 64 m channels
// take the bf16 plane kernels, the last one the one-thread-per-(point, kernel point) kernel, and the aggregate's K axis is
// laid out as [kp][64 m channels, kperm64 inside each slab] followed by [kp] for the odd channel.  This kernel splits the
// weights [K, cin, cout] into bf16 hi/lo [cout, ldk] with the same K order.
__global__ void __launch_bounds__(256) k_split_w_tail1(const float* __restrict__ w, int K, int cin, int cout, __nv_bfloat16* __restrict__ hi,
                                                       __nv_bfloat16* __restrict__ lo, int ldk)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)K * cin * cout) return;
    const int n = (int)(e % cout);
    const int kc = (int)(e / cout), k = kc / cin, c = kc - k * cin;
    const int cm = cin - 1;
    const int dst = c < cm ? k * cm + ((c & ~63) | kperm64(c & 63)) : K * cm + k;
    const float v = w[e];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[(size_t)n * ldk + dst] = h;
    lo[(size_t)n * ldk + dst] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// planes of the first `cols` channels of fp32 rows (pitch ldx) -- hi = bf16(x), lo = bf16(x - hi)
__global__ void __launch_bounds__(256) k_split_rows(const float* __restrict__ x, int ldx, int rows, int cols, __nv_bfloat16* __restrict__ hi,
                                                    __nv_bfloat16* __restrict__ lo)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)rows * cols) return;
    const int r = (int)(e / cols), c = (int)(e - (long long)r * cols);
    const float v = x[(size_t)r * ldx + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[e] = h;
    lo[e] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// first_layer_fused: measured on B200 at 1.33 ms per 32-pair step against 0.88 + 0.31 ms for aggregate_small + the CUDA-core
// contraction (one short-lived block per 32 points: latency-bound behind its staging barrier) -> opt-in until it is persistent
static int g_agg_simt = 0, g_agg_pipelined = 1, g_small_fused = 0;
// kpconv_fused (kpconv_fused.cu): 0 = never, 1 = where it wins (cin == cout == 64: the weights fit tensor memory in one
// pass), 2 = every shape it supports (cin, cout multiples of 64: S x T passes; parity tests and measurements)
static int g_fused = 1;
void kpconv_set_fused(int v) { g_fused = v; }
bool kpconv_fused_shape_ok(int64_t nq, int64_t ns, int H, int cin, int cout, int K, int ldxs);
int kpconv_fused_dev(const float* q_pts, int64_t nq, const float* s_pts, int64_t ns, const void* idx, int idx_is_i64, int H, int idx_stride,
                     const void* x_hi, const void* x_lo, int cin, int ldxs, const uint8_t* rowflag, const float* kpts, int K, float inv_extent,
                     const void* w_hi, const void* w_lo, int ldk, float* out, int cout, const int32_t* seg_starts, int nseg, double* stats_acc,
                     int64_t row0, cudaStream_t st);
void kpconv_set_small_fused(int v) { g_small_fused = v; }
void kpconv_set_agg_simt(int v) { g_agg_simt = v; }
void kpconv_set_agg_pipelined(int v) { g_agg_pipelined = v; }

template <typename IdxT, bool SPLIT>
static int launch_agg(const float* q_pts, int nq, const float* s_pts, int ns, const IdxT* idx, int H, int idx_stride, const float* x,
                      int cin, int ldx, const uint8_t* rowflag, const float* kpts, int K, float inv_extent, float* wf,
                      __nv_bfloat16* wf_hi, __nv_bfloat16* wf_lo, int ldk, float* inv_cnt, cudaStream_t st)
{
    dim3 block(AGG_WARPS * 32);
    unsigned gx = (unsigned)cdiv64(nq, AGG_WARPS);
    if (cin <= 4 && !g_agg_simt) {
#define PCRCG_AGG_SMALL(C_) k_kpconv_aggregate_small<IdxT, C_, SPLIT><<<(unsigned)cdiv64(nq, 16), 256, 0, st>>>( \
        q_pts, nq, s_pts, ns, idx, H, idx_stride, x, ldx, rowflag, kpts, K, inv_extent, wf, wf_hi, wf_lo, ldk, inv_cnt)
        if (cin == 1) PCRCG_AGG_SMALL(1);
        else if (cin == 2) PCRCG_AGG_SMALL(2);
        else if (cin == 3) PCRCG_AGG_SMALL(3);
        else PCRCG_AGG_SMALL(4);
#undef PCRCG_AGG_SMALL
        PCRCG_CUDA(cudaGetLastError());
        return PCRCG_OK;
    }
    if (cin % 8 == 0 && ns > 0 && !g_agg_simt) {
#define PCRCG_AGG_MMA(NT_) k_kpconv_aggregate_mma<IdxT, NT_, SPLIT><<<dim3(gx, (unsigned)(cin / (8 * NT_))), block, 0, st>>>( \
        q_pts, nq, s_pts, ns, idx, H, idx_stride, x, cin, ldx, rowflag, kpts, K, inv_extent, wf, wf_hi, wf_lo, ldk, inv_cnt)
        if (cin % 64 == 0) {
            PCRCG_CUDA(cudaFuncSetAttribute(k_kpconv_aggregate_mma64<IdxT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, A64_SMEM));
            k_kpconv_aggregate_mma64<IdxT, SPLIT><<<dim3((unsigned)cdiv64(nq, A64_WARPS), (unsigned)(cin / 64)), A64_WARPS * 32, A64_SMEM, st>>>(
                q_pts, nq, s_pts, ns, idx, H, idx_stride, x, cin, ldx, rowflag, kpts, K, inv_extent, wf, wf_hi, wf_lo, ldk, inv_cnt);
        }
        else if (cin % 32 == 0) PCRCG_AGG_MMA(4);
        else if (cin % 16 == 0) PCRCG_AGG_MMA(2);
        else PCRCG_AGG_MMA(1);
#undef PCRCG_AGG_MMA
        PCRCG_CUDA(cudaGetLastError());
        return PCRCG_OK;
    }
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

int gemm_tc_split_b_dev(const float* B, int ldb, int b_is_nk, int N, int K, int ldk, void* b_hi, void* b_lo, cudaStream_t st);
int gemm_tc_split_b_perm_dev(const float* B, int ldb, int b_is_nk, int N, int K, int ldk, void* b_hi, void* b_lo, int perm64, cudaStream_t st);
int gemm_tc_core_dev(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int ldk, float* C, int ldc, int M, int N, int K,
                     const float* row_scale, cudaStream_t st);
int gemm_tc_core_stats_dev(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int ldk, float* C, int ldc, int M, int N,
                           int K, const float* row_scale, cudaStream_t st, const int32_t* seg_starts, int nseg, double* stats_acc, int64_t row0);

// The [Nq, K*cin] aggregate is produced and consumed in chunks of query points through two
// alternating buffers of at most WF_CHUNK_BYTES each, which bounds the workspace for very large
// batches.  (Measured on B200: L2-sized 40 MB chunks -- meant to keep the intermediate out of HBM --
// LOSE 25 %: a chunk is then < 148 contraction tiles and both kernels run under-filled; see DESIGN.md.)
static size_t WF_CHUNK_BYTES = 8ull << 30;      // (one chunk up to 2.2 M query points at cin = 64: no split at 32 stacked pairs)

void kpconv_set_chunk_mb(int mb) { WF_CHUNK_BYTES = mb > 0 ? (size_t)mb << 20 : 8ull << 30; }      // tests: force the chunked path

static int64_t kpconv_chunk_rows(int64_t nq, size_t ldk)
{
    int64_t rows = (int64_t)(WF_CHUNK_BYTES / (ldk * sizeof(float)));
    rows = rows / 128 * 128;
    if (rows < 1024) rows = 1024;
    return rows < nq ? rows : (nq > 0 ? nq : 1);
}

size_t kpconv_ws_bytes(int64_t nq, int64_t ns, int32_t cin, int32_t K)
{
    size_t ldk = ((size_t)K * cin + 7) / 8 * 8;
    size_t chunk = (size_t)kpconv_chunk_rows(nq, ldk);
    const size_t nbuf = (int64_t)chunk < nq ? 2 : 1;          // the second (alternating) buffer only exists when the rows are chunked
    const size_t tail_planes = (cin > 64 && cin % 64 == 1) ? align_up((size_t)ns * (cin - 1) * 4, 256) : 0;      // hi + lo planes of 64 m channels
    return nbuf * align_up(chunk * ldk * sizeof(float), 256) + align_up((size_t)nq * sizeof(float), 256) + align_up((size_t)ns, 256) +
           align_up((size_t)2048 * ldk * sizeof(float), 256) + tail_planes + 2048;      // + split weights for cout <= 2048
}

// weights: [K, cin, cout] row-major (the reference's Parameter layout)
int kpconv_forward_dev(const float* q_pts, int64_t nq, const float* s_pts, int64_t ns, const void* idx, int idx_is_i64, int32_t H,
                       int32_t idx_stride, const float* x, int32_t cin, const float* kpts, int32_t K, float kp_extent,
                       const float* weights, int32_t cout, float* out, void* ws, size_t ws_bytes, cudaStream_t st,
                       const void* x_hi, const void* x_lo, int32_t ldxs, const uint8_t* rowflag_in, const int32_t* seg_starts, int32_t nseg,
                       double* stats_acc)
{
    PCRCG_REQUIRE(K >= 1 && K <= KP_MAX - 1, "kpconv: kernel_size must be in [1,15]");
    PCRCG_REQUIRE(cin >= 1 && cout >= 1 && cout <= 2048 && H >= 0 && idx_stride >= H, "kpconv: bad dimensions");
    PCRCG_REQUIRE(nq < (1ll << 31) && ns < (1ll << 31), "kpconv: too many points");
    PCRCG_REQUIRE(kp_extent > 0.f, "kpconv: KP_extent must be positive");
    PCRCG_REQUIRE((unsigned long long)ns * (unsigned long long)cin < (1ull << 32), "kpconv: feature table too large for 32-bit offsets");
    if (nq == 0) return PCRCG_OK;
    // x may be NULL when the features exist only as bf16 (hi, lo) planes with their row flags (a producer that skipped the fp32 copy)
    PCRCG_REQUIRE(x != nullptr || (x_hi != nullptr && x_lo != nullptr && rowflag_in != nullptr && cin % 64 == 0 && ldxs >= cin && ldxs % 8 == 0 &&
                                   !g_agg_simt && !gemm_force_simt_get() && gemm_tc_shape_ok((int)nq, cout, K * cin)),
                  "kpconv: fp32 features are required on this path (planes-only input needs cin %% 64 == 0, row flags and the tensor-core path)");
    const int KC = K * cin;
    const bool tc = !gemm_force_simt_get() && gemm_tc_shape_ok((int)nq, cout, KC);
    PCRCG_REQUIRE(stats_acc == nullptr || tc, "kpconv: output statistics are produced by the tensor-core contraction only");
    const int ldk = (KC + 7) / 8 * 8;
    const int64_t chunk = kpconv_chunk_rows(nq, (size_t)ldk);
    Workspace W(ws, ws_bytes);
    float* wf_buf[2];
    wf_buf[0] = W.take<float>((size_t)chunk * ldk);
    wf_buf[1] = chunk < nq ? W.take<float>((size_t)chunk * ldk) : wf_buf[0];
    float* inv_cnt = W.take<float>((size_t)nq);
    uint8_t* rowflag_ws = W.take<uint8_t>((size_t)(ns > 0 ? ns : 1));
    const uint8_t* rowflag = rowflag_in != nullptr ? rowflag_in : rowflag_ws;
    __nv_bfloat16* b_hi = (__nv_bfloat16*)W.take<float>((size_t)cout * ldk);
    __nv_bfloat16* b_lo = b_hi + (size_t)cout * ldk;
    // cin = 64 m + 1 on the tensor path: planes of the first 64 m channels are made here (see k_split_w_tail1)
    const bool tail1 = !gemm_force_simt_get() && gemm_tc_shape_ok((int)nq, cout, KC) && x != nullptr && x_hi == nullptr && cin > 64 &&
                       cin % 64 == 1 && H <= 64 && !g_agg_simt && ns > 0 && g_agg_pipelined;
    const int cin_m = cin - 1;
    __nv_bfloat16* t_hi = tail1 ? (__nv_bfloat16*)W.take<float>((size_t)ns * cin_m) : nullptr;
    __nv_bfloat16* t_lo = tail1 ? t_hi + (size_t)ns * cin_m : nullptr;
    PCRCG_REQUIRE(ws != nullptr && W.ok(), "kpconv: workspace too small (%zu < %zu)", ws_bytes, W.off);
    const float inv_extent = 1.0f / kp_extent;
    if (ns > 0 && rowflag_in == nullptr) {
        ProfScope prof(PC_KPCONV_AGG, st, 1);
        k_row_positive<<<(unsigned)cdiv64(ns, 8), 256, 0, st>>>(x, (int)ns, cin, cin, rowflag_ws);
        PCRCG_CUDA(cudaGetLastError());
    }
    // which aggregation kernel runs (the same for every chunk): bf16 planes -> ldmatrix kernels; the pipelined one writes its
    // 64-channel slabs in kperm64 order, so the weights are split with the matching K permutation
    const bool planes = tc && x_hi != nullptr && x_lo != nullptr && cin % 64 == 0 && ldxs >= cin && ldxs % 8 == 0 && !g_agg_simt && ns > 0;
    const bool pipelined = planes && g_agg_pipelined && H <= 64;
    const bool fused = pipelined && g_fused > 0 && kpconv_fused_shape_ok(nq, ns, H, cin, cout, K, ldxs) &&
                       (g_fused > 1 || (cin == 64 && cout == 64));
    if (tail1) {
        ProfScope prof(PC_KPCONV_AGG, st, 2);
        k_split_rows<<<(unsigned)cdiv64((int64_t)ns * cin_m, 256), 256, 0, st>>>(x, cin, (int)ns, cin_m, t_hi, t_lo);
        k_split_w_tail1<<<(unsigned)cdiv64((int64_t)K * cin * cout, 256), 256, 0, st>>>(weights, K, cin, cout, b_hi, b_lo, ldk);
        PCRCG_CUDA(cudaGetLastError());
    } else if (tc) {
        ProfScope prof(fused ? PC_KPCONV_FUSED : PC_GEMM, st, 0);
        PCRCG_TRY(gemm_tc_split_b_perm_dev(weights, cout, 0, cout, KC, ldk, b_hi, b_lo, pipelined ? 1 : 0, st));
    }
    if (fused) {
        // gather -> influence -> tcgen05 contraction in one kernel: the aggregate stays on the SM
        ProfScope prof(PC_KPCONV_FUSED, st, 0);
        return kpconv_fused_dev(q_pts, nq, s_pts, ns, idx, idx_is_i64, H, idx_stride, x_hi, x_lo, cin, ldxs, rowflag, kpts, K, inv_extent, b_hi,
                                b_lo, ldk, out, cout, seg_starts, nseg, stats_acc, 0, st);
    }
    const size_t idx_bytes = idx_is_i64 ? 8 : 4;
    if (cin <= 4 && !g_agg_simt && !gemm_force_simt_get() && stats_acc == nullptr && ns > 0 && g_small_fused) {
        // first layer: aggregation + contraction + 1/count in one kernel
        ProfScope prof(PC_KPCONV_AGG, st, 1);
        const bool done = idx_is_i64
            ? launch_small_fused<long long>(q_pts, (int)nq, s_pts, (int)ns, (const long long*)idx, H, idx_stride, x, cin, rowflag, kpts, K, inv_extent, weights, cout, out, st)
            : launch_small_fused<int>(q_pts, (int)nq, s_pts, (int)ns, (const int*)idx, H, idx_stride, x, cin, rowflag, kpts, K, inv_extent, weights, cout, out, st);
        if (done) {
            PCRCG_CUDA(cudaGetLastError());
            return PCRCG_OK;
        }
    }
    int it = 0;
    for (int64_t r0 = 0; r0 < nq; r0 += chunk, it++) {
        const int rows = (int)((nq - r0) < chunk ? (nq - r0) : chunk);
        float* wf = wf_buf[it & 1];
        __nv_bfloat16* wf_hi = (__nv_bfloat16*)wf;
        __nv_bfloat16* wf_lo = wf_hi + (size_t)rows * ldk;
        const float* qp = q_pts + 3 * (size_t)r0;
        const void* ip = (const char*)idx + (size_t)r0 * idx_stride * idx_bytes;
        {
            ProfScope prof(PC_KPCONV_AGG, st, 1);
            int rc;
            if (tail1) {
                // channels [0, 64 m): persistent plane kernel into K columns [kp][64 m]; channel 64 m: K columns K * 64 m + kp
                const unsigned gy = (unsigned)(cin_m / 64);
                unsigned gx = (unsigned)cdiv64(kNumSMs * 5, gy);
                if ((int64_t)gx > cdiv64(rows, AB_WARPS)) gx = (unsigned)cdiv64(rows, AB_WARPS);
                const dim3 grid(gx, gy);
                __nv_bfloat16* th = wf_hi + (size_t)K * cin_m;
                __nv_bfloat16* tl = wf_lo + (size_t)K * cin_m;
                if (idx_is_i64) {
                    k_kpconv_aggregate_bf16p<long long><<<grid, AB_WARPS * 32, ABP_SMEM, st>>>(qp, rows, s_pts, (int)ns, (const long long*)ip, H, idx_stride,
                        t_hi, t_lo, cin_m, cin_m, rowflag, kpts, K, inv_extent, wf_hi, wf_lo, ldk, inv_cnt + r0);
                    k_kpconv_aggregate_small<long long, 1, true><<<(unsigned)cdiv64(rows, 16), 256, 0, st>>>(qp, rows, s_pts, (int)ns, (const long long*)ip, H,
                        idx_stride, x + cin_m, cin, rowflag, kpts, K, inv_extent, nullptr, th, tl, ldk, inv_cnt + r0);
                } else {
                    k_kpconv_aggregate_bf16p<int><<<grid, AB_WARPS * 32, ABP_SMEM, st>>>(qp, rows, s_pts, (int)ns, (const int*)ip, H, idx_stride,
                        t_hi, t_lo, cin_m, cin_m, rowflag, kpts, K, inv_extent, wf_hi, wf_lo, ldk, inv_cnt + r0);
                    k_kpconv_aggregate_small<int, 1, true><<<(unsigned)cdiv64(rows, 16), 256, 0, st>>>(qp, rows, s_pts, (int)ns, (const int*)ip, H,
                        idx_stride, x + cin_m, cin, rowflag, kpts, K, inv_extent, nullptr, th, tl, ldk, inv_cnt + r0);
                }
                rc = cudaGetLastError() == cudaSuccess ? PCRCG_OK : PCRCG_ERR;
                if (rc) set_error("kpconv: (64 m + 1)-channel aggregate launch failed");
            } else if (pipelined) {
                // persistent: ~5 resident CTAs per SM in total, each warp walks over points with a fixed stride
                const unsigned gy = (unsigned)(cin / 64);
                unsigned gx = (unsigned)cdiv64(kNumSMs * 5, gy);
                if ((int64_t)gx > cdiv64(rows, AB_WARPS)) gx = (unsigned)cdiv64(rows, AB_WARPS);
                dim3 grid(gx, gy);
                if (idx_is_i64)
                    k_kpconv_aggregate_bf16p<long long><<<grid, AB_WARPS * 32, ABP_SMEM, st>>>(qp, rows, s_pts, (int)ns, (const long long*)ip, H, idx_stride,
                        (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, cin, ldxs, rowflag, kpts, K, inv_extent, wf_hi, wf_lo, ldk, inv_cnt + r0);
                else
                    k_kpconv_aggregate_bf16p<int><<<grid, AB_WARPS * 32, ABP_SMEM, st>>>(qp, rows, s_pts, (int)ns, (const int*)ip, H, idx_stride,
                        (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, cin, ldxs, rowflag, kpts, K, inv_extent, wf_hi, wf_lo, ldk, inv_cnt + r0);
                rc = cudaGetLastError() == cudaSuccess ? PCRCG_OK : PCRCG_ERR;
                if (rc) set_error("kpconv: pipelined bf16 aggregate launch failed");
            } else if (planes) {
                dim3 grid((unsigned)cdiv64(rows, AB_WARPS), (unsigned)(cin / 64));
                if (idx_is_i64)
                    k_kpconv_aggregate_bf16<long long><<<grid, AB_WARPS * 32, AB_SMEM, st>>>(qp, rows, s_pts, (int)ns, (const long long*)ip, H, idx_stride,
                        (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, cin, ldxs, rowflag, kpts, K, inv_extent, wf_hi, wf_lo, ldk, inv_cnt + r0);
                else
                    k_kpconv_aggregate_bf16<int><<<grid, AB_WARPS * 32, AB_SMEM, st>>>(qp, rows, s_pts, (int)ns, (const int*)ip, H, idx_stride,
                        (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, cin, ldxs, rowflag, kpts, K, inv_extent, wf_hi, wf_lo, ldk, inv_cnt + r0);
                rc = cudaGetLastError() == cudaSuccess ? PCRCG_OK : PCRCG_ERR;
                if (rc) set_error("kpconv: bf16 aggregate launch failed");
            } else if (idx_is_i64) {
                rc = tc ? launch_agg<long long, true>(qp, rows, s_pts, (int)ns, (const long long*)ip, H, idx_stride, x, cin, cin, rowflag, kpts, K, inv_extent, nullptr, wf_hi, wf_lo, ldk, inv_cnt + r0, st)
                        : launch_agg<long long, false>(qp, rows, s_pts, (int)ns, (const long long*)ip, H, idx_stride, x, cin, cin, rowflag, kpts, K, inv_extent, wf, nullptr, nullptr, ldk, inv_cnt + r0, st);
            } else {
                rc = tc ? launch_agg<int, true>(qp, rows, s_pts, (int)ns, (const int*)ip, H, idx_stride, x, cin, cin, rowflag, kpts, K, inv_extent, nullptr, wf_hi, wf_lo, ldk, inv_cnt + r0, st)
                        : launch_agg<int, false>(qp, rows, s_pts, (int)ns, (const int*)ip, H, idx_stride, x, cin, cin, rowflag, kpts, K, inv_extent, wf, nullptr, nullptr, ldk, inv_cnt + r0, st);
            }
            if (rc) return rc;
        }
        if (tc) {
            ProfScope prof(PC_GEMM, st, 0);
            PCRCG_TRY(gemm_tc_core_stats_dev(wf_hi, wf_lo, b_hi, b_lo, ldk, out + (size_t)r0 * cout, cout, rows, cout, KC, inv_cnt + r0, st,
                                             seg_starts, nseg, stats_acc, r0));
        } else {
            PCRCG_TRY(gemm_dev(wf, ldk, weights, cout, 0, out + (size_t)r0 * cout, cout, rows, cout, KC, inv_cnt + r0, st));
        }
    }
    return PCRCG_OK;
}

}  // namespace pcrcg
