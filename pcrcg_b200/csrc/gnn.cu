// Bottleneck overlap-attention GNN of PCR-CG's KPFCNN -- sm_100a.  ("next" row 3 of the scope table.)
//
// Replaces models/gcn.py (SelfAttention :98-137, MultiHeadedAttention / attention :153-172, GCN :188-217) and the
// projections / saliency scores around it (models/architectures.py:528-565).  The dense contractions of these layers go
// through the same tcgen05 / CUDA-core contraction as the encoder (gemm.cu); this file holds the rest:
//   k_knn_brute        get_graph_feature's kNN (models/gcn.py:48-51): per cloud, k+1 smallest of the reference's
//                      EXPANDED squared distance (-2 x.y + |x|^2 + |y|^2, clamped at 1e-12), first one dropped
//   k_edge_max_stats   the 1x1 conv over edge features [f_n ; f_j - f_n] is split as  W [f_n ; f_j - f_n] =
//                      (Wa - Wb) f_n + Wb f_j = u_n + v_j  (two node-level contractions instead of one over N*k edges);
//                      InstanceNorm2d + LeakyReLU are increasing per channel, so max_j act(norm(u_n + v_j)) =
//                      act(norm(u_n + max_j v_j)): this kernel produces u_n + max_j v_j and the (sum, sum of squares)
//                      of u_n + v_j over ALL edges of each cloud for the normalisation
//   k_bias_act         Conv1d bias (+ ReLU / LeakyReLU)
//   k_softmax_rows     softmax(scale * x) along rows (attention probabilities, saliency weights)
//   k_l2norm_rows      F.normalize(x, p=2, dim=1)
#include "common.cuh"

namespace pcrcg {

// ---- kNN ----------------------------------------------------------------------------------------
// One warp per query.  Keys are (distance bits << 32 | index): ascending distance, ties by index (canonical; the
// reference's torch.topk tie order is unspecified).  The k+1 smallest are extracted one after the other (each lane scans
// its stride of the cloud for the smallest key above the previous one): O(k N / 32) per query, clouds here are <= a few
// thousand coarse nodes.
__device__ __forceinline__ float knn_dist(float qx, float qy, float qz, float qn, const float* __restrict__ p)
{
    const float x = p[0], y = p[1], z = p[2];
    // models/gcn.py:28-34: dist = -2 * (src @ dst^T); dist += |src|^2; dist += |dst|^2; clamp(min=1e-12).
    // The K = 3 product is the fused chain fma(z, z', fma(y, y', x * x')) (how the reference's CPU sgemm evaluates it);
    // squared norms are sums of separately rounded squares, (x^2 + y^2) + z^2.
    const float dot = __fmaf_rn(qz, z, __fmaf_rn(qy, y, __fmul_rn(qx, x)));
    const float pn = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    float d = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), qn), pn);
    return d < 1e-12f ? 1e-12f : d;
}

__global__ void __launch_bounds__(256) k_knn_brute(const float* __restrict__ pts, int n, const int32_t* __restrict__ starts, int nb, int k,
                                                   int32_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const int c = cloud_of(starts, nb, i);
    const int s0 = starts[c], s1 = starts[c + 1];
    const float qx = pts[3 * (size_t)i], qy = pts[3 * (size_t)i + 1], qz = pts[3 * (size_t)i + 2];
    const float qn = __fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz));
    unsigned long long last = 0;
    bool have_last = false;
    for (int o = 0; o <= k; o++) {
        unsigned long long best = ~0ull;
        for (int j = s0 + lane; j < s1; j += 32) {
            const float d = knn_dist(qx, qy, qz, qn, pts + 3 * (size_t)j);
            const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
            if ((!have_last || key > last) && key < best) best = key;
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const unsigned long long x = __shfl_xor_sync(0xffffffffu, best, s);
            best = x < best ? x : best;
        }
        last = best;
        have_last = true;
        // entry 0 (the query itself in the reference's reading) is dropped; a cloud with fewer than k+1 points repeats the query
        if (o >= 1 && lane == 0) out[(size_t)i * k + (o - 1)] = best == ~0ull ? i : (int32_t)(uint32_t)(best & 0xffffffffull);
    }
}

// ---- point2node (datasets/dataloader.py:91-106) ----------------------------------------------------
// idx[i] = nearest NODE of point i inside its own cloud, under the reference's expanded squared distance
// (datasets/dataloader.py:70-90: same formula as models/gcn.py) and topk(k=1, largest=False); ties by node index.
// One warp per point; lanes stride the nodes of the cloud.  Returned indices are LOCAL to the cloud's node list, as in the
// reference (which is called once per cloud).
__global__ void __launch_bounds__(256) k_point2node(const float* __restrict__ pts, int n, const int32_t* __restrict__ pstarts,
                                                    const float* __restrict__ nodes, const int32_t* __restrict__ nstarts, int nb,
                                                    int32_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const int c = cloud_of(pstarts, nb, i);
    const int s0 = nstarts[c], s1 = nstarts[c + 1];
    const float qx = pts[3 * (size_t)i], qy = pts[3 * (size_t)i + 1], qz = pts[3 * (size_t)i + 2];
    const float qn = __fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz));
    unsigned long long best = ~0ull;
    for (int j = s0 + lane; j < s1; j += 32) {
        const float d = knn_dist(qx, qy, qz, qn, nodes + 3 * (size_t)j);
        const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(j - s0);
        best = key < best ? key : best;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const unsigned long long x = __shfl_xor_sync(0xffffffffu, best, s);
        best = x < best ? x : best;
    }
    if (lane == 0) out[i] = best == ~0ull ? 0 : (int32_t)(uint32_t)(best & 0xffffffffull);
}

// per node: number of points assigned to it and number of those flagged visible (datasets/dataloader.py:133-158)
__global__ void __launch_bounds__(256) k_node_counts(const int32_t* __restrict__ p2n, const uint8_t* __restrict__ visible, int n,
                                                     const int32_t* __restrict__ pstarts, const int32_t* __restrict__ nstarts, int nb,
                                                     int32_t* __restrict__ tot, int32_t* __restrict__ vis)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cloud_of(pstarts, nb, i);
    const int node = nstarts[c] + p2n[i];
    atomicAdd(tot + node, 1);
    if (visible[i]) atomicAdd(vis + node, 1);
}

// ---- edge max + statistics -----------------------------------------------------------------------
// One warp per node; lanes stride the channels.  m[n,c] = u[n,c] + max_j v[idx[n,j],c]; per (cloud, channel) the sum and
// the sum of squares of u[n,c] + v[idx[n,j],c] over all k edges, divided by k (so that the finaliser, which divides by
// the number of NODES of the cloud, yields the mean / variance over nodes x edges), added in fp64.
__global__ void __launch_bounds__(256) k_edge_max_stats(const float* __restrict__ u, int ldu, const float* __restrict__ v, int ldv,
                                                        const int32_t* __restrict__ idx, int n, int C, int k,
                                                        const int32_t* __restrict__ starts, int nb, float* __restrict__ m,
                                                        double* __restrict__ acc)
{
    extern __shared__ float s_part[];          // [2][C] block partial sums (the 8 nodes of a block usually share a cloud)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int node0 = blockIdx.x * 8;
    const int i = node0 + w;
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) s_part[c] = 0.f;
    __syncthreads();
    const int last_node = min(node0 + 7, n - 1);
    const int seg_first = cloud_of(starts, nb, node0), seg_last = cloud_of(starts, nb, last_node);
    const bool one_seg = seg_first == seg_last;
    if (i < n) {
        const int seg = one_seg ? seg_first : cloud_of(starts, nb, i);
        const int32_t* row = idx + (size_t)i * k;
        const float inv_k = 1.0f / (float)k;
        for (int c = lane; c < C; c += 32) {
            const float uu = u[(size_t)i * ldu + c];
            float mx = -INFINITY, s1 = 0.f, s2 = 0.f;
            for (int j = 0; j < k; j++) {
                const float e = uu + v[(size_t)row[j] * ldv + c];
                mx = fmaxf(mx, e);
                s1 += e;
                s2 = fmaf(e, e, s2);
            }
            m[(size_t)i * C + c] = mx;
            if (one_seg) {
                atomicAdd(&s_part[c], s1 * inv_k);
                atomicAdd(&s_part[C + c], s2 * inv_k);
            } else {
                atomicAdd(acc + ((size_t)seg * 2 + 0) * C + c, (double)(s1 * inv_k));
                atomicAdd(acc + ((size_t)seg * 2 + 1) * C + c, (double)(s2 * inv_k));
            }
        }
    }
    __syncthreads();
    if (one_seg)
        for (int c = threadIdx.x; c < 2 * C; c += blockDim.x)
            atomicAdd(acc + ((size_t)seg_first * 2 + (c >= C)) * C + (c >= C ? c - C : c), (double)s_part[c]);
}

// ---- small elementwise / row kernels ------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bias_act(const float* __restrict__ x, long long total, int C, const float* __restrict__ bias,
                                                  float slope, float* __restrict__ out)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    float v = x[e] + (bias != nullptr ? bias[(int)(e % C)] : 0.f);
    if (slope >= 0.f) v = v > 0.f ? v : v * slope;
    out[e] = v;
}

// x[r, 0:m] <- softmax(scale * x[r, 0:m]); one warp per row
__global__ void __launch_bounds__(256) k_softmax_rows(float* __restrict__ x, int n, int m, int ld, float scale)
{
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    float* row = x + (size_t)r * ld;
    float mx = -INFINITY;
    for (int c = lane; c < m; c += 32) mx = fmaxf(mx, row[c] * scale);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int c = lane; c < m; c += 32) {
        const float e = expf(row[c] * scale - mx);
        row[c] = e;
        sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    for (int c = lane; c < m; c += 32) row[c] *= inv;
}

// out[r,:] = x[r,:] / max(|x[r,:]|_2, eps)      (F.normalize, models/architectures.py:543)
__global__ void __launch_bounds__(256) k_l2norm_rows(const float* __restrict__ x, int n, int C, float eps, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    const float* row = x + (size_t)r * C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) ss = fmaf(row[c], row[c], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
    for (int c = lane; c < C; c += 32) out[(size_t)r * C + c] = row[c] * inv;
}

// ---- host side ---------------------------------------------------------------------------------
int knn_dev(const float* pts, int64_t n, const int32_t* cloud_starts, int32_t nb, int32_t k, int32_t* out, cudaStream_t st)
{
    PCRCG_REQUIRE(n >= 0 && n < (1ll << 31) && nb >= 1 && k >= 1 && k <= 64, "knn: bad dimensions");
    if (n == 0) return PCRCG_OK;
    ProfScope prof(PC_RADIUS_QUERY, st, 1);
    k_knn_brute<<<(unsigned)cdiv64(n, 8), 256, 0, st>>>(pts, (int)n, cloud_starts, nb, k, out);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int point2node_dev(const float* pts, int64_t n, const int32_t* pstarts, const float* nodes, const int32_t* nstarts, int32_t nb, int32_t* out,
                   cudaStream_t st)
{
    PCRCG_REQUIRE(n >= 0 && n < (1ll << 31) && nb >= 1, "point2node: bad dimensions");
    if (n == 0) return PCRCG_OK;
    ProfScope prof(PC_RADIUS_QUERY, st, 1);
    k_point2node<<<(unsigned)cdiv64(n, 8), 256, 0, st>>>(pts, (int)n, pstarts, nodes, nstarts, nb, out);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int node_counts_dev(const int32_t* p2n, const uint8_t* visible, int64_t n, const int32_t* pstarts, const int32_t* nstarts, int32_t nb,
                    int32_t* tot, int32_t* vis, cudaStream_t st)
{
    PCRCG_REQUIRE(n >= 0 && n < (1ll << 31) && nb >= 1, "node_counts: bad dimensions");
    if (n == 0) return PCRCG_OK;
    ProfScope prof(PC_POOL, st, 1);
    k_node_counts<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(p2n, visible, (int)n, pstarts, nstarts, nb, tot, vis);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int edge_max_stats_dev(const float* u, int32_t ldu, const float* v, int32_t ldv, const int32_t* idx, int64_t n, int32_t C, int32_t k,
                       const int32_t* cloud_starts, int32_t nb, float* m, double* stats_acc, cudaStream_t st)
{
    PCRCG_REQUIRE(n >= 0 && n < (1ll << 31) && C >= 1 && C <= 4096 && k >= 1 && nb >= 1, "edge_max_stats: bad dimensions");
    if (n == 0) return PCRCG_OK;
    ProfScope prof(PC_POOL, st, 1);
    k_edge_max_stats<<<(unsigned)cdiv64(n, 8), 256, 2 * (size_t)C * sizeof(float), st>>>(u, ldu, v, ldv, idx, (int)n, C, k, cloud_starts, nb, m,
                                                                                      stats_acc);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int bias_act_dev(const float* x, int64_t n, int32_t C, const float* bias, float slope, float* out, cudaStream_t st)
{
    PCRCG_REQUIRE(n >= 0 && C >= 1, "bias_act: bad dimensions");
    if (n == 0) return PCRCG_OK;
    ProfScope prof(PC_NORM, st, 1);
    const long long total = (long long)n * C;
    k_bias_act<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(x, total, C, bias, slope, out);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int softmax_rows_dev(float* x, int64_t n, int32_t m, int32_t ld, float scale, cudaStream_t st)
{
    PCRCG_REQUIRE(n >= 0 && n < (1ll << 31) && m >= 1 && ld >= m, "softmax_rows: bad dimensions");
    if (n == 0) return PCRCG_OK;
    ProfScope prof(PC_NORM, st, 1);
    k_softmax_rows<<<(unsigned)cdiv64(n, 8), 256, 0, st>>>(x, (int)n, m, ld, scale);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int l2norm_rows_dev(const float* x, int64_t n, int32_t C, float eps, float* out, cudaStream_t st)
{
    PCRCG_REQUIRE(n >= 0 && n < (1ll << 31) && C >= 1, "l2norm_rows: bad dimensions");
    if (n == 0) return PCRCG_OK;
    ProfScope prof(PC_NORM, st, 1);
    k_l2norm_rows<<<(unsigned)cdiv64(n, 8), 256, 0, st>>>(x, (int)n, C, eps, out);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

}  // namespace pcrcg
