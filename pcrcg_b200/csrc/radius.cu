// Batched fixed-radius neighbour search on a hashed uniform grid -- sm_100a.
//
// Replaces cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:211-332 (batch_nanoflann_neighbors)
// and the nanoflann kd-tree behind it.  Contract:
//   * neighbour set  : { j in same cloud : ((dx*dx + dy*dy) + dz*dz) < radius*radius }, fp32, no FMA
//   * order          : ascending (d2, index)  -- the project's canonical tie rule
//   * rows           : global support indices, padded with the shadow index Ns, truncated to `width`
//
// Grid: per cloud, cell edge cs = radius*(1+1e-3) + 8*ulp(max|coord|) so that any pair accepted by
// the fp32 distance test lies in adjacent cells despite rounding in the cell computation.  Occupied
// cells live in an open-addressing table (2*Ns+1 slots per cloud); supports are bucketed by
// table slot (counting sort: per-slot count, exclusive scan, atomic cursor) so each cell is one
// contiguous run of (x,y,z,index) float4 records.
//
// Search kernel: one warp per query.  Lanes 0..26 resolve the 27 adjacent cells, the warp then
// walks the concatenated candidate runs 32 at a time, tests d2 and compacts hits with
// ballot/popc into a per-warp shared-memory list of 64-bit keys (d2 bits << 32 | index), sorts it
// (bitonic, in shared memory) and writes the row.
#include "common.cuh"

namespace pcrcg {

constexpr uint32_t EMPTY = 0xffffffffu;
constexpr int RQ_WARPS = 8;
constexpr int RQ_CAND = 384;         // candidate record indices staged per warp (longer candidate sets use the search path)
constexpr int RQ_CAP = 256;          // matches kept in shared memory per warp; larger rows take the slow path

// ---- bbox (same encoding as subsample.cu; kept local so both TUs stay self-contained) ------------
__global__ void k_rbbox_init(int* __restrict__ bbox, int nb)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb * 6) bbox[i] = (i % 6) < 3 ? 0x7fffffff : (int)0x80000000;
}

__global__ void __launch_bounds__(256) k_rbbox(const float* __restrict__ pts, int n, const int32_t* __restrict__ starts, int nb,
                                               int* __restrict__ bbox)
{
    bbox_accumulate(pts, n, starts, nb, bbox);
}

// per cloud: grid origin (support bbox min) and cell edge
__global__ void k_grid_meta(const int* __restrict__ bbox, int nb, float radius, float4* __restrict__ meta)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nb) return;
    float mn[3], amax = 0.f;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        mn[d] = ord2f(bbox[6 * c + d]);
        float mx = ord2f(bbox[6 * c + 3 + d]);
        amax = fmaxf(amax, fmaxf(fabsf(mn[d]), fabsf(mx)));
    }
    float cs = radius * 1.001f + 8.0f * (amax * 1.1920929e-7f);
    meta[c] = make_float4(mn[0], mn[1], mn[2], cs);
}

__device__ __forceinline__ int cell_coord(float x, float o, float cs)
{
    float f = floorf((x - o) / cs);
    f = fminf(fmaxf(f, -2.0f), 2097150.0f);
    return (int)f;
}

__device__ __forceinline__ uint32_t cell_slot(uint64_t key, uint32_t cap)
{
    return __umulhi((uint32_t)(mix64(key) >> 32), cap);      // uniform in [0, cap)
}

__device__ __forceinline__ uint64_t cell_key(int cx, int cy, int cz)
{
    return (uint64_t)(uint32_t)cx | ((uint64_t)(uint32_t)cy << 21) | ((uint64_t)(uint32_t)cz << 42);
}

__global__ void __launch_bounds__(256) k_cell_keys(const float* __restrict__ s, int ns, const int32_t* __restrict__ sstarts, int nb,
                                                   const float4* __restrict__ meta, uint64_t* __restrict__ keys)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    int c = cloud_of(sstarts, nb, i);
    float4 m = meta[c];
    int cx = cell_coord(s[3 * (size_t)i], m.x, m.w), cy = cell_coord(s[3 * (size_t)i + 1], m.y, m.w),
        cz = cell_coord(s[3 * (size_t)i + 2], m.z, m.w);
    keys[i] = cell_key(max(cx, 0), max(cy, 0), max(cz, 0));
}

__global__ void __launch_bounds__(256) k_cell_insert(const uint64_t* __restrict__ keys, int ns, const int32_t* __restrict__ sstarts,
                                                     int nb, uint32_t* __restrict__ rep, uint32_t* __restrict__ slot,
                                                     uint32_t* __restrict__ cell_count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    int c = cloud_of(sstarts, nb, i);
    int s0 = sstarts[c], len = sstarts[c + 1] - s0;
    uint32_t cap = 2u * (uint32_t)len + 1u, toff = 2u * (uint32_t)s0 + (uint32_t)c;
    uint64_t key = keys[i];
    uint32_t h = cell_slot(key, cap);
    while (true) {
        uint32_t old = atomicCAS(&rep[toff + h], EMPTY, (uint32_t)i);
        if (old == EMPTY || keys[old] == key) break;
        h = h + 1u == cap ? 0u : h + 1u;
    }
    slot[i] = toff + h;
    atomicAdd(&cell_count[toff + h], 1u);
}

// Counting sort by cell slot: cell_start = exclusive scan of cell_count; every point takes the next free
// position of its cell.  The order INSIDE a cell is arbitrary (atomic cursor) -- harmless: rows are sorted by
// (d2, index) afterwards, so results do not depend on it.
__global__ void __launch_bounds__(256) k_cell_scatter(const float* __restrict__ s, const uint32_t* __restrict__ slot, int ns,
                                                      const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cursor,
                                                      float4* __restrict__ rec, uint2* __restrict__ range)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    const uint32_t sl = slot[i];
    const uint32_t st = cell_start[sl];
    const uint32_t k = atomicAdd(&cursor[sl], 1u);
    rec[st + k] = make_float4(s[3 * (size_t)i], s[3 * (size_t)i + 1], s[3 * (size_t)i + 2], __uint_as_float((uint32_t)i));
    if (k == 0) range[sl] = make_uint2(st, cell_start[sl + 1]);
}

__device__ __forceinline__ float d2_ref(float qx, float qy, float qz, float4 p)
{
    float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Bitonic sort of R*32 64-bit keys held R per lane (element e = r*32 + lane), ascending, in registers.
template <int R>
__device__ __forceinline__ void warp_bitonic(unsigned long long (&v)[R], int lane)
{
#pragma unroll
    for (int k = 2; k <= R * 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if ((r & jr) == 0) {
                        const bool asc = ((r * 32) & k) == 0;             // compile-time after unrolling
                        const unsigned long long a = v[r], b = v[r | jr];
                        const bool swap = (a < b) != asc;
                        v[r] = swap ? b : a;
                        v[r | jr] = swap ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[r], j);
                    // this lane keeps the smaller key iff (lower half of the pair) == (ascending run); keys are distinct, so
                    // "take the partner's key" = (mine < partner's) != keep_small: one 64-bit compare and one select per exchange
                    const bool keep_small = ((lane & j) == 0) == ((((r * 32) | lane) & k) == 0);
                    v[r] = ((v[r] < o) != keep_small) ? o : v[r];
                }
            }
        }
    }
}

template <int R>
__device__ __forceinline__ void sort_and_write(const unsigned long long* buf, int nm, int lane, int32_t* row, int width, int ns)
{
    unsigned long long v[R];
#pragma unroll
    for (int r = 0; r < R; r++) { int e = r * 32 + lane; v[r] = e < nm ? buf[e] : ~0ull; }
    warp_bitonic<R>(v, lane);
#pragma unroll
    for (int r = 0; r < R; r++) { int e = r * 32 + lane; if (e < width) row[e] = e < nm ? (int32_t)(uint32_t)(v[r] & 0xffffffffull) : ns; }
    for (int e = R * 32 + lane; e < width; e += 32) row[e] = ns;
}

// One warp per query.  rows == nullptr: count only.
__global__ void __launch_bounds__(RQ_WARPS * 32) k_radius_query(
    const float* __restrict__ q, int nq, const int32_t* __restrict__ qstarts, const int32_t* __restrict__ sstarts, int nb,
    const float4* __restrict__ meta, const uint64_t* __restrict__ skeys, const uint32_t* __restrict__ rep,
    const uint2* __restrict__ range, const float4* __restrict__ rec, float r2, int ns, int width, int row_stride,
    int32_t* __restrict__ rows, int32_t* __restrict__ counts, int32_t* __restrict__ maxcount)
{
    __shared__ unsigned long long s_buf[RQ_WARPS][RQ_CAP];
    __shared__ uint32_t s_pre[RQ_WARPS][28];
    __shared__ uint32_t s_st[RQ_WARPS][27];
    __shared__ uint32_t s_cand[RQ_WARPS][RQ_CAND];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * RQ_WARPS + w;
    if (i >= nq) return;                      // warp-uniform
    unsigned long long* buf = s_buf[w];

    const int c = cloud_of(qstarts, nb, i);
    const float qx = q[3 * (size_t)i], qy = q[3 * (size_t)i + 1], qz = q[3 * (size_t)i + 2];
    const float4 m = meta[c];
    const int s0 = sstarts[c], slen = sstarts[c + 1] - s0;
    uint32_t st = 0, cnt = 0;
    if (lane < 27 && slen > 0) {
        int cx = cell_coord(qx, m.x, m.w) + (lane % 3) - 1;
        int cy = cell_coord(qy, m.y, m.w) + ((lane / 3) % 3) - 1;
        int cz = cell_coord(qz, m.z, m.w) + (lane / 9) - 1;
        if (cx >= 0 && cy >= 0 && cz >= 0) {
            uint64_t key = cell_key(cx, cy, cz);
            uint32_t cap = 2u * (uint32_t)slen + 1u, toff = 2u * (uint32_t)s0 + (uint32_t)c;
            uint32_t h = cell_slot(key, cap);
            while (true) {
                uint32_t r = rep[toff + h];
                if (r == EMPTY) break;
                if (skeys[r] == key) { uint2 rg = range[toff + h]; st = rg.x; cnt = rg.y - rg.x; break; }
                h = h + 1u == cap ? 0u : h + 1u;
            }
        }
    }
    // exclusive prefix of the 27 run lengths
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    if (lane < 27) { s_pre[w][lane] = inc - cnt; s_st[w][lane] = st; }
    if (lane == 26) s_pre[w][27] = inc;
    __syncwarp();
    const uint32_t C = s_pre[w][27];

    int nm = 0;                                // matches so far (warp-uniform)
    if (C <= (uint32_t)RQ_CAND) {
        // expand the 27 runs into a flat list of record indices (each lane expands its own run)
        if (lane < 27) {
            const uint32_t p0 = s_pre[w][lane];
            for (uint32_t t = 0; t < cnt; t++) s_cand[w][p0 + t] = st + t;
        }
        __syncwarp();
        for (uint32_t base = 0; base < C; base += 32) {
            const uint32_t t = base + lane;
            bool hit = false;
            unsigned long long key = 0;
            if (t < C) {
                const float4 p = rec[s_cand[w][t]];
                const float d2 = d2_ref(qx, qy, qz, p);
                hit = d2 < r2;
                key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(p.w);
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const int pos = nm + __popc(bal & ((1u << lane) - 1u));
                if (pos < RQ_CAP) buf[pos] = key;
            }
            nm += __popc(bal);
        }
    } else {
        for (uint32_t base = 0; base < C; base += 32) {
            uint32_t t = base + lane;
            bool hit = false;
            unsigned long long key = 0;
            if (t < C) {
                int lo = 0, hi = 27;               // find run: largest k with pre[k] <= t
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pre[w][mid] <= t) lo = mid; else hi = mid; }
                float4 p = rec[s_st[w][lo] + (t - s_pre[w][lo])];
                float d2 = d2_ref(qx, qy, qz, p);
                hit = d2 < r2;
                key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(p.w);
            }
            uint32_t bal = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                int pos = nm + __popc(bal & ((1u << lane) - 1u));
                if (pos < RQ_CAP) buf[pos] = key;
            }
            nm += __popc(bal);
        }
    }
    if (lane == 0) {
        if (counts) counts[i] = nm;
        if (maxcount) atomicMax(maxcount, nm);
    }
    if (rows == nullptr) return;
    int32_t* row = rows + (size_t)i * row_stride;
    __syncwarp();

    if (nm <= 32) {
        sort_and_write<1>(buf, nm, lane, row, width, ns);
    } else if (nm <= 64) {
        sort_and_write<2>(buf, nm, lane, row, width, ns);
    } else if (nm <= 128) {
        sort_and_write<4>(buf, nm, lane, row, width, ns);
    } else if (nm <= RQ_CAP) {
        int mp = 1;
        while (mp < nm) mp <<= 1;
        for (int k = nm + lane; k < mp; k += 32) buf[k] = ~0ull;
        __syncwarp();
        for (int k = 2; k <= mp; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int x = lane; x < mp; x += 32) {
                    int y = x ^ j;
                    if (y > x) {
                        unsigned long long a = buf[x], b = buf[y];
                        bool asc = (x & k) == 0;
                        if ((a > b) == asc) { buf[x] = b; buf[y] = a; }
                    }
                }
                __syncwarp();
            }
        }
        for (int k = lane; k < width; k += 32) row[k] = k < nm ? (int32_t)(uint32_t)(buf[k] & 0xffffffffull) : ns;
    } else {
        // slow path (row longer than the shared list): repeated extraction of the next smallest key
        unsigned long long last = 0;
        bool have_last = false;
        const int nout = nm < width ? nm : width;
        for (int o = 0; o < nout; o++) {
            unsigned long long best = ~0ull;
            for (uint32_t base = 0; base < C; base += 32) {
                uint32_t t = base + lane;
                if (t < C) {
                    int lo = 0, hi = 27;
                    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pre[w][mid] <= t) lo = mid; else hi = mid; }
                    float4 p = rec[s_st[w][lo] + (t - s_pre[w][lo])];
                    float d2 = d2_ref(qx, qy, qz, p);
                    unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(p.w);
                    if (d2 < r2 && (!have_last || key > last) && key < best) best = key;
                }
            }
#pragma unroll
            for (int of = 16; of > 0; of >>= 1) {
                unsigned long long x = __shfl_xor_sync(0xffffffffu, best, of);
                best = x < best ? x : best;
            }
            last = best;
            have_last = true;
            if (lane == 0) row[o] = (int32_t)(uint32_t)(best & 0xffffffffull);
        }
        for (int k = nout + lane; k < width; k += 32) row[k] = ns;
    }
}

// ------------------------------------------------------------------------------------------------
struct RadWS {
    int32_t *qstarts, *sstarts; int* bbox; float4* meta; uint64_t* keys; uint32_t* rep; uint2* range; uint32_t* slot;
    uint32_t* iota; uint32_t* sslot; uint32_t* sidx; float4* rec; void* prim; size_t prim_bytes;
};

static size_t rad_layout(Workspace& W, int64_t ns, int32_t nb, RadWS* o)
{
    RadWS r;
    size_t n1 = (size_t)(ns > 0 ? ns : 1);
    r.qstarts = W.take<int32_t>(nb + 1);
    r.sstarts = W.take<int32_t>(nb + 1);
    r.bbox = W.take<int>((size_t)nb * 6);
    r.meta = W.take<float4>(nb);
    r.keys = W.take<uint64_t>(n1);
    r.rep = W.take<uint32_t>(2 * n1 + nb);
    r.range = W.take<uint2>(2 * n1 + nb);
    r.slot = W.take<uint32_t>(n1);
    r.iota = W.take<uint32_t>(2 * n1 + nb + 1);        // per-slot cursor
    r.sslot = W.take<uint32_t>(2 * n1 + nb + 2);       // per-slot count -> start (exclusive scan, +1 total)
    r.sidx = W.take<uint32_t>(1);
    r.rec = W.take<float4>(n1);
    r.prim_bytes = scan_ws_bytes(2 * (int64_t)n1 + nb + 1);
    r.prim = W.take<char>(r.prim_bytes);
    if (o) *o = r;
    return W.off;
}

size_t radius_ws_bytes(int64_t nq, int64_t ns, int32_t nb)
{
    (void)nq;
    Workspace W(nullptr, 0);
    return rad_layout(W, ns, nb, nullptr) + 256;
}

// Builds the support grid into the workspace (state lives entirely in ws; a later
// radius_query_dev call with the same ws reuses it).
int radius_build_dev(const float* s, int64_t ns, const int32_t* s_lens, int32_t nb, float radius, void* ws, size_t ws_bytes,
                     cudaStream_t st)
{
    PCRCG_REQUIRE(ns >= 0 && ns < (1ll << 30), "radius search: Ns out of range");
    PCRCG_REQUIRE(nb >= 1 && nb < 65536, "radius search: number of clouds out of range");
    PCRCG_REQUIRE(radius > 0.f, "radius search: radius must be positive");
    Workspace W(ws, ws_bytes);
    RadWS r;
    rad_layout(W, ns, nb, &r);
    PCRCG_REQUIRE(ws != nullptr && W.ok(), "radius search: workspace too small (%zu < %zu)", ws_bytes, W.off);
    const int NS = (int)ns;
    const unsigned gs = (unsigned)cdiv64(NS > 0 ? NS : 1, 256);
    ProfScope prof(PC_RADIUS_BUILD, st, 11);
    const size_t nslots = 2 * (size_t)NS + nb;
    PCRCG_TRY(cloud_starts(s_lens, nb, r.sstarts, st));
    k_rbbox_init<<<(nb * 6 + 255) / 256, 256, 0, st>>>(r.bbox, nb);
    k_rbbox<<<gs, 256, 0, st>>>(s, NS, r.sstarts, nb, r.bbox);
    k_grid_meta<<<(nb + 127) / 128, 128, 0, st>>>(r.bbox, nb, radius, r.meta);
    k_cell_keys<<<gs, 256, 0, st>>>(s, NS, r.sstarts, nb, r.meta, r.keys);
    PCRCG_CUDA(cudaMemsetAsync(r.rep, 0xff, sizeof(uint32_t) * nslots, st));
    PCRCG_CUDA(cudaMemsetAsync(r.iota, 0, sizeof(uint32_t) * (nslots + 1), st));
    PCRCG_CUDA(cudaMemsetAsync(r.sslot, 0, sizeof(uint32_t) * (nslots + 2), st));
    k_cell_insert<<<gs, 256, 0, st>>>(r.keys, NS, r.sstarts, nb, r.rep, r.slot, r.sslot);
    PCRCG_CUDA(cudaGetLastError());
    PCRCG_TRY(exclusive_scan_u32(r.sslot, r.sslot, (int64_t)nslots, r.prim, r.prim_bytes, st));
    k_cell_scatter<<<gs, 256, 0, st>>>(s, r.slot, NS, r.sslot, r.iota, r.rec, r.range);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int radius_query_dev(const float* q, int64_t nq, const int32_t* q_lens, int64_t ns, int32_t nb, float radius, int32_t width,
                     int32_t row_stride, int32_t* rows, int32_t* counts, int32_t* maxcount, void* ws, size_t ws_bytes,
                     cudaStream_t st)
{
    PCRCG_REQUIRE(nq >= 0 && nq < (1ll << 30), "radius search: Nq out of range");
    Workspace W(ws, ws_bytes);
    RadWS r;
    rad_layout(W, ns, nb, &r);
    PCRCG_REQUIRE(ws != nullptr && W.ok(), "radius search: workspace too small");
    PCRCG_REQUIRE(rows == nullptr || (width >= 0 && row_stride >= width), "radius search: bad row geometry");
    if (nq == 0) return PCRCG_OK;
    ProfScope prof(PC_RADIUS_QUERY, st, 3);
    PCRCG_TRY(cloud_starts(q_lens, nb, r.qstarts, st));
    if (maxcount) PCRCG_CUDA(cudaMemsetAsync(maxcount, 0, sizeof(int32_t), st));
    const float r2 = radius * radius;      // neighbors.cpp:226 (fp32 product)
    k_radius_query<<<(unsigned)cdiv64(nq, RQ_WARPS), RQ_WARPS * 32, 0, st>>>(
        q, (int)nq, r.qstarts, r.sstarts, nb, r.meta, r.keys, r.rep, r.range, r.rec, r2, (int)ns, width, row_stride, rows,
        counts, maxcount);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

}  // namespace pcrcg
