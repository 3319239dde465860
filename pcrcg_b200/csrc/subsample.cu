// Grid (voxel) barycentre subsampling of a stacked batch of clouds -- sm_100a.
//
// Replaces  cpp_wrappers.zip!cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-211
// (grid_subsampling / batch_grid_subsampling).  Bit-exact contract:
//   * voxel key   : fp32 origin = floor(min*(1/dl))*dl, true fp32 divides, size_t key arithmetic
//   * barycentre  : fp32 sums in ORIGINAL point order, times (float)(1.0/(double)count)
//   * output order: iteration order of libstdc++'s unordered_map<size_t,...>, reproduced in closed
//                   form per rehash epoch (k_order), clouds processed independently.
//
// Pipeline (all stream ordered, no host sync):
//   bbox -> origin/NX/NY -> keys -> hash insert (voxel slot per point) -> stable radix sort by slot
//   -> run heads (= first occurrence + count) -> scan (first-occurrence rank) -> sequential
//   barycentres -> per-cloud order emulation -> gather to output.
// Optional per-point features / integer classes (grid_subsampling.cpp:34-102; not used by the KPConv pyramid,
// datasets/dataloader.py:289 passes neither): subsample_batch_ex_dev runs the same pipeline and then, from the sorted runs and
// the final lists it leaves in the workspace, the per-voxel feature means (k_bary_feat), the class votes (k_label_vote,
// label_vote.h) and their gather into the output order (k_gather_extra).
#include "common.cuh"
#include "subsample_extras.h"

namespace pcrcg {

constexpr uint32_t EMPTY = 0xffffffffu;

// bucket-count schedule of libstdc++ (GCC 13) unordered_map, max_load_factor 1 (see oracle/port.c)
__constant__ uint32_t c_sched[27] = { 13u, 29u, 59u, 127u, 257u, 541u, 1109u, 2357u, 5087u, 10273u, 20753u, 42043u,
                                      85229u, 172933u, 351061u, 712697u, 1447153u, 2938679u, 5967347u, 12117689u,
                                      24607243u, 49969847u, 101473717u, 206062531u, 418450807u, 849747061u, 1725587117u };

// ------------------------------------------------------------------------------------------------
__global__ void k_bbox_init(int* __restrict__ bbox, int nb)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb * 6) bbox[i] = (i % 6) < 3 ? 0x7fffffff : (int)0x80000000;
}

// bbox[c] = {min x,y,z, max x,y,z} in the ordered-int encoding
__global__ void __launch_bounds__(256) k_bbox(const float* __restrict__ pts, int n, const int32_t* __restrict__ starts, int nb,
                                              int* __restrict__ bbox)
{
    bbox_accumulate(pts, n, starts, nb, bbox);
}

__device__ __forceinline__ uint64_t f2size_t(float f) { return (uint64_t)(long long)f; }

// grid_subsampling.cpp:25-31
__global__ void k_origin(const int* __restrict__ bbox, int nb, float dl, float* __restrict__ origin, uint64_t* __restrict__ nxny)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nb) return;
    float inv = __fdiv_rn(1.0f, dl);
    float o[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        float mn = ord2f(bbox[6 * c + d]);
        o[d] = __fmul_rn(floorf(__fmul_rn(mn, inv)), dl);
        origin[3 * c + d] = o[d];
    }
    float mxx = ord2f(bbox[6 * c + 3]), mxy = ord2f(bbox[6 * c + 4]);
    nxny[2 * c + 0] = f2size_t(floorf(__fdiv_rn(__fsub_rn(mxx, o[0]), dl))) + 1ull;
    nxny[2 * c + 1] = f2size_t(floorf(__fdiv_rn(__fsub_rn(mxy, o[1]), dl))) + 1ull;
}

// grid_subsampling.cpp:53-56
__global__ void __launch_bounds__(256) k_keys(const float* __restrict__ pts, int n, const int32_t* __restrict__ starts, int nb,
                                              float dl, const float* __restrict__ origin, const uint64_t* __restrict__ nxny,
                                              uint64_t* __restrict__ keys)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cloud_of(starts, nb, i);
    float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
    uint64_t iX = f2size_t(floorf(__fdiv_rn(__fsub_rn(x, origin[3 * c + 0]), dl)));
    uint64_t iY = f2size_t(floorf(__fdiv_rn(__fsub_rn(y, origin[3 * c + 1]), dl)));
    uint64_t iZ = f2size_t(floorf(__fdiv_rn(__fsub_rn(z, origin[3 * c + 2]), dl)));
    uint64_t NX = nxny[2 * c], NY = nxny[2 * c + 1];
    keys[i] = iX + NX * iY + NX * NY * iZ;
}

// Open-addressing insert into the cloud's table region [2*start_c + c, +2*len_c+1).  A slot stores the
// index of one representative point; key equality is tested through keys[representative].
__global__ void __launch_bounds__(256) k_insert(const uint64_t* __restrict__ keys, int n, const int32_t* __restrict__ starts, int nb,
                                                uint32_t* __restrict__ rep, uint32_t* __restrict__ slot, uint32_t* __restrict__ iota)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cloud_of(starts, nb, i);
    int s0 = starts[c], len = starts[c + 1] - s0;
    uint32_t cap = 2u * (uint32_t)len + 1u;
    uint32_t toff = 2u * (uint32_t)s0 + (uint32_t)c;
    uint64_t key = keys[i];
    uint32_t h = __umulhi((uint32_t)(mix64(key) >> 32), cap);
    while (true) {
        uint32_t old = atomicCAS(&rep[toff + h], EMPTY, (uint32_t)i);
        if (old == EMPTY || keys[old] == key) break;
        h = h + 1u == cap ? 0u : h + 1u;
    }
    slot[i] = toff + h;
    iota[i] = (uint32_t)i;
}

// run heads in the slot-sorted order: head's point index is the voxel's first occurrence
__global__ void __launch_bounds__(256) k_heads(const uint32_t* __restrict__ sslot, const uint32_t* __restrict__ sidx, int n,
                                               uint32_t* __restrict__ flag)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (j == 0 || sslot[j] != sslot[j - 1]) flag[sidx[j]] = 1u;
}

// grid_subsampling.h:74-79 + .cpp:87 : sequential fp32 sum in original order, * (float)(1.0/count)
__global__ void __launch_bounds__(256) k_bary(const float* __restrict__ pts, const uint64_t* __restrict__ keys,
                                              const uint32_t* __restrict__ sslot, const uint32_t* __restrict__ sidx, int n,
                                              const uint32_t* __restrict__ rank, float* __restrict__ baryU, uint64_t* __restrict__ keyU)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t s = sslot[j];
    if (j != 0 && sslot[j - 1] == s) return;
    uint32_t first = sidx[j];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    int cnt = 0;
    for (int t = j; t < n && sslot[t] == s; t++) {
        size_t p = sidx[t];
        sx = __fadd_rn(sx, pts[3 * p]);
        sy = __fadd_rn(sy, pts[3 * p + 1]);
        sz = __fadd_rn(sz, pts[3 * p + 2]);
        cnt++;
    }
    float a = __double2float_rn(1.0 / (double)cnt);
    size_t u = rank[first];
    baryU[3 * u + 0] = __fmul_rn(sx, a);
    baryU[3 * u + 1] = __fmul_rn(sy, a);
    baryU[3 * u + 2] = __fmul_rn(sz, a);
    keyU[u] = keys[first];
}

// grid_subsampling.cpp:181-204 : keep the head of each cloud's list when max_p > 0
__global__ void k_outlens(const uint32_t* __restrict__ rank, const int32_t* __restrict__ starts, int nb, int max_p,
                          int32_t* __restrict__ out_lens)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nb) return;
    int m = (int)(rank[starts[c + 1]] - rank[starts[c]]);
    out_lens[c] = (max_p > 0 && m > max_p) ? max_p : m;
}

// ------------------------------------------------------------------------------------------------
// Order emulation: one CTA per cloud.  Elements are the cloud's voxels named by first-occurrence
// rank e in [0,M).  Epoch j inserts ranks [done, min(M, P_j)) after rehashing the current list to
// P_j buckets.  One epoch over sequence S (= current list followed by the new ranks):
//   buckets b(e) = key % P;  groups (buckets) in REVERSE order of first appearance in S,
//   elements of a group in REVERSE order of appearance  (SURVEY.md App. A.2).
// out position of S[p] = (sum of sizes of groups whose first appearance f' > f(p)) + #{q in group: q > p}.
constexpr int ORD_THREADS = 1024;

__global__ void __launch_bounds__(ORD_THREADS) k_order(const uint64_t* __restrict__ keyU, const uint32_t* __restrict__ rank,
                                                       const int32_t* __restrict__ starts, const int32_t* __restrict__ out_lens,
                                                       const int32_t* __restrict__ out_base, const float* __restrict__ baryU,
                                                       float* __restrict__ out_pts, uint32_t* __restrict__ seqA,
                                                       uint32_t* __restrict__ seqB, uint32_t* __restrict__ nxt,
                                                       uint32_t* __restrict__ aux, uint32_t* __restrict__ rr,
                                                       uint32_t* __restrict__ ff, uint32_t* __restrict__ head_all)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int c = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int s0 = starts[c];
    const uint32_t Ub = rank[s0];
    const int M = (int)(rank[starts[c + 1]] - Ub);
    const uint64_t* key = keyU + Ub;
    uint32_t* L = seqA + Ub;
    uint32_t* Lo = seqB + Ub;
    nxt += Ub; aux += Ub; rr += Ub; ff += Ub;
    uint32_t* head = head_all + (9ull * (uint64_t)s0) / 4ull + 32ull * (uint64_t)c;

    int done = 0;
    for (int j = 0; done < M; j++) {
        const uint32_t P = c_sched[j];
        const int hi = (uint32_t)M < P ? M : (int)P;
        for (int p = done + t; p < hi; p += ORD_THREADS) L[p] = (uint32_t)p;
        for (uint32_t b = t; b < P; b += ORD_THREADS) head[b] = EMPTY;
        __syncthreads();
        for (int p = t; p < hi; p += ORD_THREADS) {
            uint32_t b = (uint32_t)(key[L[p]] % (uint64_t)P);
            rr[p] = b;
            nxt[p] = atomicExch(&head[b], (uint32_t)p);
        }
        __syncthreads();
        for (int p = t; p < hi; p += ORD_THREADS) {
            uint32_t b = rr[p];
            uint32_t f = EMPTY, cnt = 0, r = 0;
            for (uint32_t q = head[b]; q != EMPTY; q = nxt[q]) {
                f = q < f ? q : f;
                cnt++;
                r += q > (uint32_t)p;
            }
            aux[p] = f == (uint32_t)p ? cnt : 0u;
            rr[p] = r;
            ff[p] = f;
        }
        __syncthreads();
        // suffix-exclusive scan of aux over [0,hi): walk positions from the back
        if (t == 0) carry_s = 0;
        __syncthreads();
        for (int base = 0; base < hi; base += ORD_THREADS) {
            int i = base + t;                 // reversed index
            int p = hi - 1 - i;
            uint32_t v = i < hi ? aux[p] : 0u;
            uint32_t inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += x;
            }
            if (lane == 31) warp_tot[w] = inc;
            __syncthreads();
            uint32_t woff = 0, total = 0;
#pragma unroll
            for (int k = 0; k < 32; k++) { uint32_t x = warp_tot[k]; if (k < w) woff += x; total += x; }
            uint32_t carry = carry_s;
            if (i < hi) aux[p] = carry + woff + inc - v;
            __syncthreads();
            if (t == 0) carry_s = carry + total;
            __syncthreads();
        }
        for (int p = t; p < hi; p += ORD_THREADS) Lo[aux[ff[p]] + rr[p]] = L[p];
        __syncthreads();
        uint32_t* tmp = L; L = Lo; Lo = tmp;
        done = hi;
    }
    const int m_out = out_lens[c];
    const size_t ob = (size_t)out_base[c];
    for (int k = t; k < m_out * 3; k += ORD_THREADS) {
        int e = k / 3, d = k - 3 * e;
        out_pts[(ob + e) * 3 + d] = baryU[((size_t)Ub + L[e]) * 3 + d];
    }
}

// ------------------------------------------------------------------------------------------------
// Feature means, class votes and their gather into the output order (thread bodies: subsample_extras.h)
__global__ void __launch_bounds__(256) k_bary_feat(const float* __restrict__ feat, int fdim, const uint32_t* __restrict__ sslot,
                                                   const uint32_t* __restrict__ sidx, int n, const uint32_t* __restrict__ rank,
                                                   float* __restrict__ featU)
{
    bary_feat_thread((int)(blockIdx.x * blockDim.x + threadIdx.x), (int)blockIdx.y, (int)gridDim.y, feat, fdim, sslot, sidx, n, rank, featU);
}

// *status = 1 when a voxel holds more distinct labels than the order model covers
__global__ void __launch_bounds__(128) k_label_vote(const int32_t* __restrict__ cls, int ldim, const uint32_t* __restrict__ sslot,
                                                    const uint32_t* __restrict__ sidx, int n, const uint32_t* __restrict__ rank,
                                                    int32_t* __restrict__ clsU, int32_t* __restrict__ status)
{
    if (label_vote_thread((int)(blockIdx.x * blockDim.x + threadIdx.x), (int)blockIdx.y, (int)gridDim.y, cls, ldim, sslot, sidx, n, rank, clsU))
        atomicExch(status, 1);
}

__global__ void __launch_bounds__(256) k_gather_extra(const uint32_t* __restrict__ rank, const int32_t* __restrict__ starts,
                                                      const int32_t* __restrict__ out_lens, const int32_t* __restrict__ out_base,
                                                      const uint32_t* __restrict__ seqA, const uint32_t* __restrict__ seqB,
                                                      const float* __restrict__ featU, int fdim, float* __restrict__ out_feat,
                                                      const int32_t* __restrict__ clsU, int ldim, int32_t* __restrict__ out_cls)
{
    gather_extra_thread((int)blockIdx.x, (int)threadIdx.x, (int)blockDim.x, rank, starts, out_lens, out_base, seqA, seqB, c_sched, featU,
                        fdim, out_feat, clsU, ldim, out_cls);
}

// ------------------------------------------------------------------------------------------------
struct SubWS {
    int32_t* starts; int* bbox; float* origin; uint64_t* nxny; uint64_t* keys; uint32_t* rep; uint32_t* slot;
    uint32_t* iota; uint32_t* sslot; uint32_t* sidx; uint32_t* rank; float* baryU; uint64_t* keyU; int32_t* out_base;
    uint32_t *seqA, *seqB, *nxt, *aux, *rr, *ff, *head;
    void* prim; size_t prim_bytes;
};

static size_t sub_layout(Workspace& W, int64_t n, int32_t nb, SubWS* o)
{
    SubWS s;
    size_t n1 = (size_t)(n > 0 ? n : 1);
    s.starts = W.take<int32_t>(nb + 1);
    s.bbox = W.take<int>((size_t)nb * 6);
    s.origin = W.take<float>((size_t)nb * 3);
    s.nxny = W.take<uint64_t>((size_t)nb * 2);
    s.keys = W.take<uint64_t>(n1);
    s.rep = W.take<uint32_t>(2 * n1 + nb);
    s.slot = W.take<uint32_t>(n1);
    s.iota = W.take<uint32_t>(n1);
    s.sslot = W.take<uint32_t>(n1);
    s.sidx = W.take<uint32_t>(n1);
    s.rank = W.take<uint32_t>(n1 + 1);
    s.baryU = W.take<float>(3 * n1);
    s.keyU = W.take<uint64_t>(n1);
    s.out_base = W.take<int32_t>(nb + 1);
    s.seqA = W.take<uint32_t>(n1); s.seqB = W.take<uint32_t>(n1); s.nxt = W.take<uint32_t>(n1);
    s.aux = W.take<uint32_t>(n1); s.rr = W.take<uint32_t>(n1); s.ff = W.take<uint32_t>(n1);
    s.head = W.take<uint32_t>((9 * n1) / 4 + 32 * (size_t)nb + 64);
    size_t pb = sort_ws_bytes(n) > scan_ws_bytes(n) ? sort_ws_bytes(n) : scan_ws_bytes(n);
    s.prim = W.take<char>(pb);
    s.prim_bytes = pb;
    if (o) *o = s;
    return W.off;
}

size_t subsample_ws_bytes(int64_t n, int32_t nb)
{
    Workspace W(nullptr, 0);
    return sub_layout(W, n, nb, nullptr) + 256;
}

int subsample_batch_dev(const float* pts, int64_t n, const int32_t* lens, int32_t nb, float dl, int32_t max_p,
                        float* out_pts, int32_t* out_lens, void* ws, size_t ws_bytes, cudaStream_t st)
{
    PCRCG_REQUIRE(n >= 0 && n < (1ll << 30), "subsample: n out of range");
    PCRCG_REQUIRE(nb >= 1 && nb < 65536, "subsample: number of clouds out of range");
    PCRCG_REQUIRE(dl > 0.f, "subsample: sampleDl must be positive");
    Workspace W(ws, ws_bytes);
    SubWS s;
    sub_layout(W, n, nb, &s);
    PCRCG_REQUIRE(ws != nullptr && W.ok(), "subsample: workspace too small (%zu < %zu)", ws_bytes, W.off);
    const int N = (int)n;
    const unsigned gb = (unsigned)cdiv64(N > 0 ? N : 1, 256);

    int nbits = 1;
    while ((1ull << nbits) < 2ull * (uint64_t)N + (uint64_t)nb + 1ull) nbits++;
    ProfScope prof(PC_SUBSAMPLE, st, 15 + 5 * ((nbits + 7) / 8));
    PCRCG_TRY(cloud_starts(lens, nb, s.starts, st));
    k_bbox_init<<<(nb * 6 + 255) / 256, 256, 0, st>>>(s.bbox, nb);
    k_bbox<<<gb, 256, 0, st>>>(pts, N, s.starts, nb, s.bbox);
    k_origin<<<(nb + 127) / 128, 128, 0, st>>>(s.bbox, nb, dl, s.origin, s.nxny);
    k_keys<<<gb, 256, 0, st>>>(pts, N, s.starts, nb, dl, s.origin, s.nxny, s.keys);
    PCRCG_CUDA(cudaMemsetAsync(s.rep, 0xff, sizeof(uint32_t) * (2 * (size_t)N + nb), st));
    k_insert<<<gb, 256, 0, st>>>(s.keys, N, s.starts, nb, s.rep, s.slot, s.iota);
    PCRCG_CUDA(cudaGetLastError());
    PCRCG_TRY(radix_sort_pairs(s.slot, s.iota, s.sslot, s.sidx, N, nbits, s.prim, s.prim_bytes, st));
    PCRCG_CUDA(cudaMemsetAsync(s.rank, 0, sizeof(uint32_t) * ((size_t)N + 1), st));
    k_heads<<<gb, 256, 0, st>>>(s.sslot, s.sidx, N, s.rank);
    PCRCG_CUDA(cudaGetLastError());
    PCRCG_TRY(exclusive_scan_u32(s.rank, s.rank, N, s.prim, s.prim_bytes, st));
    k_bary<<<gb, 256, 0, st>>>(pts, s.keys, s.sslot, s.sidx, N, s.rank, s.baryU, s.keyU);
    k_outlens<<<(nb + 127) / 128, 128, 0, st>>>(s.rank, s.starts, nb, max_p, out_lens);
    PCRCG_CUDA(cudaGetLastError());
    PCRCG_TRY(cloud_starts(out_lens, nb, s.out_base, st));
    k_order<<<nb, ORD_THREADS, 0, st>>>(s.keyU, s.rank, s.starts, out_lens, s.out_base, s.baryU, out_pts,
                                        s.seqA, s.seqB, s.nxt, s.aux, s.rr, s.ff, s.head);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

// Workspace of the variant with features / classes: the plain layout first (so that subsample_batch_dev finds its buffers where
// it always does), then the per-voxel feature means and votes in first-occurrence order.
static size_t sub_layout_ex(Workspace& W, int64_t n, int32_t nb, int32_t fdim, int32_t ldim, SubWS* o, float** featU, int32_t** clsU)
{
    sub_layout(W, n, nb, o);
    size_t n1 = (size_t)(n > 0 ? n : 1);
    float* f = W.take<float>(n1 * (size_t)(fdim > 0 ? fdim : 0) + 1);
    int32_t* c = W.take<int32_t>(n1 * (size_t)(ldim > 0 ? ldim : 0) + 1);
    if (featU) *featU = f;
    if (clsU) *clsU = c;
    return W.off;
}

size_t subsample_ex_ws_bytes(int64_t n, int32_t nb, int32_t fdim, int32_t ldim)
{
    Workspace W(nullptr, 0);
    return sub_layout_ex(W, n, nb, fdim, ldim, nullptr, nullptr, nullptr) + 256;
}

// features [n, fdim] / classes [n, ldim] may each be nullptr (then its dim is ignored).  out_features [n, fdim], out_classes
// [n, ldim] (upper bounds, like out_pts).  status (device int32, required with classes): set to 1 when some voxel holds more than
// LV_CAP distinct labels in one column (its vote is then not the reference's); the caller reads it when it next synchronises.
int subsample_batch_ex_dev(const float* pts, int64_t n, const int32_t* lens, int32_t nb, float dl, int32_t max_p, const float* features,
                           int32_t fdim, const int32_t* classes, int32_t ldim, float* out_pts, int32_t* out_lens, float* out_features,
                           int32_t* out_classes, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t st)
{
    PCRCG_REQUIRE(features == nullptr || (fdim >= 1 && fdim <= 65535 && out_features != nullptr), "subsample: features need 1 <= fdim <= 65535 and an output");
    PCRCG_REQUIRE(classes == nullptr || (ldim >= 1 && ldim <= 65535 && out_classes != nullptr && status != nullptr),
                  "subsample: classes need 1 <= ldim <= 65535, an output and a status word");
    // grid_subsampling.cpp:157-158 slices the classes of every cloud after the first with a wrong end offset when ldim > 1
    // (reads out of bounds): there is no reference behaviour to reproduce
    PCRCG_REQUIRE(classes == nullptr || ldim == 1 || nb == 1, "subsample: classes with more than one column are defined for a single cloud only "
                  "(the reference mis-slices them for later clouds, grid_subsampling.cpp:157-158)");
    if (features == nullptr) fdim = 0;
    if (classes == nullptr) ldim = 0;
    Workspace W(ws, ws_bytes);
    SubWS s;
    float* featU = nullptr;
    int32_t* clsU = nullptr;
    sub_layout_ex(W, n, nb, fdim, ldim, &s, &featU, &clsU);
    PCRCG_REQUIRE(ws != nullptr && W.ok(), "subsample: workspace too small (%zu < %zu)", ws_bytes, W.off);
    PCRCG_TRY(subsample_batch_dev(pts, n, lens, nb, dl, max_p, out_pts, out_lens, ws, ws_bytes, st));
    if (fdim == 0 && ldim == 0) return PCRCG_OK;
    const int N = (int)n;
    count_launches((fdim ? 1 : 0) + (ldim ? 1 : 0) + 1);
    if (fdim) {
        const dim3 g((unsigned)cdiv64(N > 0 ? N : 1, 256), (unsigned)(fdim < 64 ? fdim : 64));
        k_bary_feat<<<g, 256, 0, st>>>(features, fdim, s.sslot, s.sidx, N, s.rank, featU);
    }
    if (ldim) {
        PCRCG_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
        const dim3 g((unsigned)cdiv64(N > 0 ? N : 1, 128), (unsigned)(ldim < 64 ? ldim : 64));
        k_label_vote<<<g, 128, 0, st>>>(classes, ldim, s.sslot, s.sidx, N, s.rank, clsU, status);
    }
    k_gather_extra<<<nb, 256, 0, st>>>(s.rank, s.starts, out_lens, s.out_base, s.seqA, s.seqB, featU, fdim, fdim ? out_features : nullptr,
                                       clsU, ldim, ldim ? out_classes : nullptr);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

}  // namespace pcrcg
