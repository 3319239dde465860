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
constexpr int RQ_QCHUNK = 16;        // queries of one cell handled by one work unit of the cell-centric kernel

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

// koff = 0: support grid (coordinates are >= 0 inside the supports' own bounding box); koff = 2: a QUERY set binned into the
// support grid's cells (queries may lie up to two cells below the origin: cell_coord clamps there)
__global__ void __launch_bounds__(256) k_cell_keys(const float* __restrict__ s, int ns, const int32_t* __restrict__ sstarts, int nb,
                                                   const float4* __restrict__ meta, uint64_t* __restrict__ keys, int koff)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    int c = cloud_of(sstarts, nb, i);
    float4 m = meta[c];
    int cx = cell_coord(s[3 * (size_t)i], m.x, m.w), cy = cell_coord(s[3 * (size_t)i + 1], m.y, m.w),
        cz = cell_coord(s[3 * (size_t)i + 2], m.z, m.w);
    keys[i] = koff ? cell_key(cx + koff, cy + koff, cz + koff) : cell_key(max(cx, 0), max(cy, 0), max(cz, 0));
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
                                                      float4* __restrict__ rec, uint2* __restrict__ range,
                                                      const int32_t* __restrict__ starts, int nb, uint2* __restrict__ cells,
                                                      uint32_t* __restrict__ ncells)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t sl = 0, nu = 0, cl = 0;
    if (i < ns) {
        sl = slot[i];
        const uint32_t st = cell_start[sl];
        const uint32_t k = atomicAdd(&cursor[sl], 1u);
        rec[st + k] = make_float4(s[3 * (size_t)i], s[3 * (size_t)i + 1], s[3 * (size_t)i + 2], __uint_as_float((uint32_t)i));
        if (k == 0) {
            const uint32_t en = cell_start[sl + 1];
            range[sl] = make_uint2(st, en);
            // work units of the cell-centric search: (cell, chunk of RQ_QCHUNK of its points), in arbitrary order.  A unit packs
            // the cloud (16 bits) with the chunk number; cells of a query set binned into a coarser grid hold ~50 points, and
            // one warp per such cell would leave most of the machine idle at the small pyramid levels.
            nu = (en - st + RQ_QCHUNK - 1) / RQ_QCHUNK;
            cl = (uint32_t)cloud_of(starts, nb, i);
        }
    }
    // one atomic per warp: exclusive prefix of the lanes' unit counts (a plain atomicAdd of a per-lane amount on one
    // address is not warp-aggregated by the compiler: measured 183 us against 72 us for the 1.25 M-point grid)
    const int lane = threadIdx.x & 31;
    uint32_t inc = nu;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    uint32_t base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(ncells, total);
    base = __shfl_sync(0xffffffffu, base, 31) + inc - nu;
    for (uint32_t u = 0; u < nu; u++) cells[base + u] = make_uint2(sl, cl | (u << 16));
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


// Bitonic sort of R*32 32-bit keys held R per lane (element e = r*32 + lane), ascending, in registers: one shuffle and one
// predicated min/max per exchange (the 64-bit (d2, index) network costs two of each plus a two-instruction compare).
template <int R>
__device__ __forceinline__ void warp_bitonic32(uint32_t (&v)[R], int lane)
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
                        const bool asc = ((r * 32) & k) == 0;
                        const uint32_t a = v[r], b = v[r | jr];
                        v[r] = asc ? min(a, b) : max(a, b);
                        v[r | jr] = asc ? max(a, b) : min(a, b);
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const uint32_t o = __shfl_xor_sync(0xffffffffu, v[r], j);
                    const bool keep_small = ((lane & j) == 0) == ((((r * 32) | lane) & k) == 0);
                    v[r] = keep_small ? min(v[r], o) : max(v[r], o);
                }
            }
        }
    }
}

// Staged cells: the hits of a query are sorted by a 32-bit key = (upper bits of the fp32 d2 pattern | position t of the
// candidate in the staged list, TB bits), then the exact 64-bit (d2, index) keys are rebuilt from the staged records, written
// to buf in that order and VERIFIED: truncating d2 is monotone, so the order can only be wrong between neighbours whose d2
// agree in the kept bits -- an adjacent out-of-order pair.  If one exists (exact ties on lattice data, else ~1e-4 of the
// rows) odd-even transposition passes on buf finish the job.  Result: exactly the ascending (d2, index) order.
template <int R, int TB>
__device__ __forceinline__ void sort32_and_write(const uint32_t* key32, unsigned long long* buf, const float4* cand, float qx, float qy,
                                                 float qz, int nm, int lane, int32_t* row, int width, int ns)
{
    uint32_t v[R];
#pragma unroll
    for (int r = 0; r < R; r++) { const int e = r * 32 + lane; v[r] = e < nm ? key32[e] : 0xffffffffu; }
    warp_bitonic32<R>(v, lane);
    __syncwarp();                                               // key32 aliases buf: every lane has read its keys
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int e = r * 32 + lane;
        if (e < nm) {
            const float4 p = cand[v[r] & ((1u << TB) - 1u)];
            buf[e] = ((unsigned long long)__float_as_uint(d2_ref(qx, qy, qz, p)) << 32) | (unsigned long long)__float_as_uint(p.w);
        }
    }
    __syncwarp();
    bool bad = false;
#pragma unroll
    for (int r = 0; r < R; r++) { const int e = r * 32 + lane; if (e + 1 < nm) bad |= buf[e] > buf[e + 1]; }
    if (__any_sync(0xffffffffu, bad)) {
        bool swapped = true;
        while (swapped) {                                       // warp-uniform
            swapped = false;
            for (int ph = 0; ph < 2; ph++) {
                for (int e = ph + 2 * lane; e + 1 < nm; e += 64) {
                    const unsigned long long a = buf[e], b = buf[e + 1];
                    if (a > b) { buf[e] = b; buf[e + 1] = a; swapped = true; }
                }
                __syncwarp();
            }
            swapped = __any_sync(0xffffffffu, swapped);
        }
    }
    for (int e = lane; e < width; e += 32) row[e] = e < nm ? (int32_t)(uint32_t)(buf[e] & 0xffffffffull) : ns;
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
// CELL-CENTRIC search: one warp per occupied QUERY cell (claimed from a device counter).  The 27 adjacent support cells are
// resolved ONCE per cell and their records staged ONCE in shared memory (padded to a multiple of 32 with far-away points);
// every query of the cell then tests the staged candidates (LDS.128, exact no-FMA d2, ballot/popc compaction of 64-bit
// (d2, index) keys) and sorts / writes its row as in k_radius_query.  Against the one-warp-per-query kernel this removes,
// per query, 27 dependent hash probes, the run expansion and ~110 scattered global record loads.
// A query set that is not the support set is first binned into the support grid's cells (radius_query_cells_dev).
// Cells with more than CAND candidates, and rows with more than CAP hits, take the search / extraction paths of the
// per-query kernel (same code), so capacity never limits correctness.
template <int CAND, int CAP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_radius_cells(
    const uint2* __restrict__ cells, const uint32_t* __restrict__ ncells_p, uint32_t* __restrict__ claim, const uint64_t* __restrict__ qkeys,
    const uint32_t* __restrict__ qrep, const uint2* __restrict__ qrange, const float4* __restrict__ qrec, int koff,
    const int32_t* __restrict__ sstarts, const uint64_t* __restrict__ skeys, const uint32_t* __restrict__ rep,
    const uint2* __restrict__ range, const float4* __restrict__ rec, float r2, int ns, int width, int row_stride,
    int32_t* __restrict__ rows, int32_t* __restrict__ counts, int32_t* __restrict__ maxcount)
{
    extern __shared__ __align__(16) uint8_t smem_rc[];
    constexpr int WARP_BYTES = CAND * 16 + CAP * 8 + 64 * 4;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float4* cand = reinterpret_cast<float4*>(smem_rc + (size_t)w * WARP_BYTES);
    unsigned long long* buf = reinterpret_cast<unsigned long long*>(cand + CAND);
    uint32_t* key32 = reinterpret_cast<uint32_t*>(buf) + CAP;      // 32-bit sort keys of a staged cell: the upper half of buf
    constexpr int TB = CAND <= 512 ? 9 : 10;                       // bits of a candidate position
    constexpr uint32_t TBMASK = (1u << TB) - 1u;
    static_assert(CAND <= (1 << TB), "candidate positions must fit the sort key");
    uint32_t* s_pre = reinterpret_cast<uint32_t*>(buf + CAP);       // [28]
    uint32_t* s_st = s_pre + 28;                                    // [27]
    const uint32_t ncells = *ncells_p;
    int wmax = 0;

    while (true) {
        uint32_t cell = 0;
        if (lane == 0) cell = atomicAdd(claim, 1u);
        cell = __shfl_sync(0xffffffffu, cell, 0);
        if (cell >= ncells) break;
        const uint2 ce = cells[cell];                               // (table slot of the query cell, cloud | chunk << 16)
        uint2 qr = qrange[ce.x];
        qr.x += (ce.y >> 16) * RQ_QCHUNK;
        qr.y = min(qr.y, qr.x + RQ_QCHUNK);
        const uint64_t qkey = qkeys[qrep[ce.x]];
        const int ccx = (int)(qkey & 0x1fffffu) - koff, ccy = (int)((qkey >> 21) & 0x1fffffu) - koff, ccz = (int)((qkey >> 42) & 0x1fffffu) - koff;
        const int c = (int)(ce.y & 0xffffu);
        const int s0 = sstarts[c], slen = sstarts[c + 1] - s0;
        uint32_t st = 0, cnt = 0;
        if (lane < 27 && slen > 0) {
            const int cx = ccx + (lane % 3) - 1, cy = ccy + ((lane / 3) % 3) - 1, cz = ccz + (lane / 9) - 1;
            if (cx >= 0 && cy >= 0 && cz >= 0) {
                const uint64_t key = cell_key(cx, cy, cz);
                const uint32_t cap = 2u * (uint32_t)slen + 1u, toff = 2u * (uint32_t)s0 + (uint32_t)c;
                uint32_t h = cell_slot(key, cap);
                while (true) {
                    const uint32_t r = rep[toff + h];
                    if (r == EMPTY) break;
                    if (skeys[r] == key) { const uint2 rg = range[toff + h]; st = rg.x; cnt = rg.y - rg.x; break; }
                    h = h + 1u == cap ? 0u : h + 1u;
                }
            }
        }
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += x;
        }
        __syncwarp();                                               // the previous cell's readers of s_pre / cand are done
        if (lane < 27) { s_pre[lane] = inc - cnt; s_st[lane] = st; }
        if (lane == 26) s_pre[27] = inc;
        const uint32_t C = __shfl_sync(0xffffffffu, inc, 26);
        const bool staged = C <= (uint32_t)CAND;
        const uint32_t Cpad = (C + 31u) & ~31u;
        if (staged) {
            if (lane < 27) {
                const uint32_t p0 = inc - cnt;
                for (uint32_t t = 0; t < cnt; t++) cand[p0 + t] = rec[st + t];
            }
            if (C + lane < Cpad) cand[C + lane] = make_float4(1e30f, 1e30f, 1e30f, 0.f);      // never a hit
        }
        __syncwarp();

        float4 qnext = qrec[qr.x];                                  // a unit holds >= 1 query; the next record is requested one query ahead
        for (uint32_t qi = qr.x; qi < qr.y; qi++) {
            const float4 qp = qnext;
            if (qi + 1 < qr.y) qnext = qrec[qi + 1];
            const float qx = qp.x, qy = qp.y, qz = qp.z;
            const int i = (int)__float_as_uint(qp.w);
            int nm = 0;
            if (staged) {
                for (uint32_t base = 0; base < Cpad; base += 32) {
                    const float4 p = cand[base + lane];
                    const float d2 = d2_ref(qx, qy, qz, p);
                    const bool hit = d2 < r2;
                    const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                    if (hit) {
                        const int pos = nm + __popc(bal & ((1u << lane) - 1u));
                        if (pos < CAP) key32[pos] = (__float_as_uint(d2) & ~(uint32_t)(TBMASK)) | (base + lane);
                    }
                    nm += __popc(bal);
                }
            } else {
                for (uint32_t base = 0; base < C; base += 32) {
                    const uint32_t t = base + lane;
                    bool hit = false;
                    unsigned long long key = 0;
                    if (t < C) {
                        int lo = 0, hi = 27;
                        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_pre[mid] <= t) lo = mid; else hi = mid; }
                        const float4 p = rec[s_st[lo] + (t - s_pre[lo])];
                        const float d2 = d2_ref(qx, qy, qz, p);
                        hit = d2 < r2;
                        key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(p.w);
                    }
                    const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                    if (hit) {
                        const int pos = nm + __popc(bal & ((1u << lane) - 1u));
                        if (pos < CAP) buf[pos] = key;
                    }
                    nm += __popc(bal);
                }
            }
            if (lane == 0 && counts) counts[i] = nm;
            wmax = max(wmax, nm);
            if (rows != nullptr) {
                int32_t* row = rows + (size_t)i * row_stride;
                __syncwarp();
                if (staged && nm <= 32) {
                    sort32_and_write<1, TB>(key32, buf, cand, qx, qy, qz, nm, lane, row, width, ns);
                } else if (staged && nm <= 64) {
                    sort32_and_write<2, TB>(key32, buf, cand, qx, qy, qz, nm, lane, row, width, ns);
                } else if (staged && nm <= 128 && CAP >= 128) {
                    sort32_and_write<4, TB>(key32, buf, cand, qx, qy, qz, nm, lane, row, width, ns);
                } else if (staged && nm <= CAP) {
                    // wide rows of a staged cell: exact keys first, then the shared-memory bitonic sort below
                    for (int b = 0; b < nm; b += 32) {              // warp-uniform trip count
                        const int e = b + lane;
                        unsigned long long kk = 0;
                        if (e < nm) {
                            const float4 p = cand[key32[e] & TBMASK];
                            kk = ((unsigned long long)__float_as_uint(d2_ref(qx, qy, qz, p)) << 32) | (unsigned long long)__float_as_uint(p.w);
                        }
                        __syncwarp();                               // key32 aliases the upper half of buf (entries >= 2b - CAP)
                        if (e < nm) buf[e] = kk;
                        __syncwarp();
                    }
                }
                if (staged && nm <= 128 && CAP >= 128) {
                    // written above
                } else if (!staged && nm <= 32) {
                    sort_and_write<1>(buf, nm, lane, row, width, ns);
                } else if (!staged && nm <= 64) {
                    sort_and_write<2>(buf, nm, lane, row, width, ns);
                } else if (!staged && nm <= 128 && CAP >= 128) {
                    sort_and_write<4>(buf, nm, lane, row, width, ns);
                } else if (nm <= CAP) {
                    int mp = 1;
                    while (mp < nm) mp <<= 1;
                    for (int k = nm + lane; k < mp; k += 32) buf[k] = ~0ull;
                    __syncwarp();
                    for (int k = 2; k <= mp; k <<= 1) {
                        for (int j = k >> 1; j > 0; j >>= 1) {
                            for (int x = lane; x < mp; x += 32) {
                                const int y = x ^ j;
                                if (y > x) {
                                    const unsigned long long a = buf[x], b = buf[y];
                                    const bool asc = (x & k) == 0;
                                    if ((a > b) == asc) { buf[x] = b; buf[y] = a; }
                                }
                            }
                            __syncwarp();
                        }
                    }
                    for (int k = lane; k < width; k += 32) row[k] = k < nm ? (int32_t)(uint32_t)(buf[k] & 0xffffffffull) : ns;
                } else {
                    // row longer than the shared list: repeated extraction of the next smallest key (reads the runs again)
                    unsigned long long last = 0;
                    bool have_last = false;
                    const int nout = nm < width ? nm : width;
                    for (int o = 0; o < nout; o++) {
                        unsigned long long best = ~0ull;
                        for (uint32_t base = 0; base < C; base += 32) {
                            const uint32_t t = base + lane;
                            if (t < C) {
                                int lo = 0, hi = 27;
                                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_pre[mid] <= t) lo = mid; else hi = mid; }
                                const float4 p = rec[s_st[lo] + (t - s_pre[lo])];
                                const float d2 = d2_ref(qx, qy, qz, p);
                                const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(p.w);
                                if (d2 < r2 && (!have_last || key > last) && key < best) best = key;
                            }
                        }
#pragma unroll
                        for (int of = 16; of > 0; of >>= 1) {
                            const unsigned long long x = __shfl_xor_sync(0xffffffffu, best, of);
                            best = x < best ? x : best;
                        }
                        last = best;
                        have_last = true;
                        if (lane == 0) row[o] = (int32_t)(uint32_t)(best & 0xffffffffull);
                    }
                    for (int k = nout + lane; k < width; k += 32) row[k] = ns;
                }
                __syncwarp();                                       // buf is rewritten by the next query
            }
        }
    }
    if (lane == 0 && maxcount && wmax > 0) atomicMax(maxcount, wmax);
}

// ------------------------------------------------------------------------------------------------
// A binned point set: hash table of occupied cells (per cloud), records sorted by cell, list of occupied cells.
struct GridWS {
    int32_t* starts; uint64_t* keys; uint32_t* rep; uint2* range; uint32_t* slot; uint32_t* cursor; uint32_t* sslot; float4* rec;
    uint2* cells; uint32_t* counters;       // counters[0] = number of occupied cells, counters[1] = claim cursor of the search kernel
    void* prim; size_t prim_bytes;
};
struct RadWS {
    int32_t* qstarts; int* bbox; float4* meta; GridWS g;
};

static void grid_layout(Workspace& W, int64_t n, int32_t nb, GridWS* o)
{
    GridWS r;
    const size_t n1 = (size_t)(n > 0 ? n : 1);
    r.starts = W.take<int32_t>(nb + 1);
    r.keys = W.take<uint64_t>(n1);
    r.rep = W.take<uint32_t>(2 * n1 + nb);
    r.range = W.take<uint2>(2 * n1 + nb);
    r.slot = W.take<uint32_t>(n1);
    r.cursor = W.take<uint32_t>(2 * n1 + nb + 1);      // per-slot cursor
    r.sslot = W.take<uint32_t>(2 * n1 + nb + 2);       // per-slot count -> start (exclusive scan, +1 total)
    r.rec = W.take<float4>(n1);
    r.cells = W.take<uint2>(n1);
    r.counters = W.take<uint32_t>(4);
    r.prim_bytes = scan_ws_bytes(2 * (int64_t)n1 + nb + 1);
    r.prim = W.take<char>(r.prim_bytes);
    if (o) *o = r;
}

static size_t rad_layout(Workspace& W, int64_t ns, int32_t nb, RadWS* o)
{
    RadWS r;
    r.qstarts = W.take<int32_t>(nb + 1);
    r.bbox = W.take<int>((size_t)nb * 6);
    r.meta = W.take<float4>(nb);
    grid_layout(W, ns, nb, &r.g);
    if (o) *o = r;
    return W.off;
}

size_t radius_ws_bytes(int64_t nq, int64_t ns, int32_t nb)
{
    (void)nq;
    Workspace W(nullptr, 0);
    return rad_layout(W, ns, nb, nullptr) + 256;
}

// workspace of a QUERY set binned into an existing support grid (radius_query_cells_dev)
size_t radius_query_ws_bytes(int64_t nq, int32_t nb)
{
    Workspace W(nullptr, 0);
    grid_layout(W, nq, nb, nullptr);
    return W.off + 256;
}

// bins `pts` (stacked clouds, lens) into the cells defined by meta (origin, edge per cloud)
static int grid_build(const float* pts, int n, const int32_t* lens, int32_t nb, const float4* meta, int koff, const GridWS& g, cudaStream_t st)
{
    const unsigned gs = (unsigned)cdiv64(n > 0 ? n : 1, 256);
    const size_t nslots = 2 * (size_t)n + nb;
    PCRCG_TRY(cloud_starts(lens, nb, g.starts, st));
    k_cell_keys<<<gs, 256, 0, st>>>(pts, n, g.starts, nb, meta, g.keys, koff);
    PCRCG_CUDA(cudaMemsetAsync(g.rep, 0xff, sizeof(uint32_t) * nslots, st));
    PCRCG_CUDA(cudaMemsetAsync(g.cursor, 0, sizeof(uint32_t) * (nslots + 1), st));
    PCRCG_CUDA(cudaMemsetAsync(g.sslot, 0, sizeof(uint32_t) * (nslots + 2), st));
    PCRCG_CUDA(cudaMemsetAsync(g.counters, 0, sizeof(uint32_t) * 4, st));
    k_cell_insert<<<gs, 256, 0, st>>>(g.keys, n, g.starts, nb, g.rep, g.slot, g.sslot);
    PCRCG_CUDA(cudaGetLastError());
    PCRCG_TRY(exclusive_scan_u32(g.sslot, g.sslot, (int64_t)nslots, g.prim, g.prim_bytes, st));
    k_cell_scatter<<<gs, 256, 0, st>>>(pts, g.slot, n, g.sslot, g.cursor, g.rec, g.range, g.starts, nb, g.cells, g.counters);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
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
    PCRCG_TRY(cloud_starts(s_lens, nb, r.g.starts, st));
    k_rbbox_init<<<(nb * 6 + 255) / 256, 256, 0, st>>>(r.bbox, nb);
    k_rbbox<<<gs, 256, 0, st>>>(s, NS, r.g.starts, nb, r.bbox);
    k_grid_meta<<<(nb + 127) / 128, 128, 0, st>>>(r.bbox, nb, radius, r.meta);
    return grid_build(s, NS, s_lens, nb, r.meta, 0, r.g, st);
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
        q, (int)nq, r.qstarts, r.g.starts, nb, r.meta, r.g.keys, r.g.rep, r.g.range, r.g.rec, r2, (int)ns, width, row_stride, rows,
        counts, maxcount);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

template <int CAND, int CAP, int WARPS>
static int launch_cells(const GridWS& qg, int koff, const RadWS& r, float r2, int ns, int width, int row_stride, int32_t* rows, int32_t* counts,
                        int32_t* maxcount, int64_t nq, cudaStream_t st)
{
    constexpr int SMEM = WARPS * (CAND * 16 + CAP * 8 + 64 * 4);
    auto kern = k_radius_cells<CAND, CAP, WARPS>;
    PCRCG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    int per_sm = (227 * 1024) / (SMEM + 1024);
    const int by_threads = 2048 / (WARPS * 32);
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)kNumSMs * per_sm;
    const int64_t need = cdiv64(nq, WARPS);                    // never more warps than queries (a cell holds >= 1 query)
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WARPS * 32, SMEM, st>>>(qg.cells, qg.counters, qg.counters + 1, qg.keys, qg.rep, qg.range, qg.rec, koff, r.g.starts,
                                                   r.g.keys, r.g.rep, r.g.range, r.g.rec, r2, ns, width, row_stride, rows, counts, maxcount);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

// Cell-centric search.  queries_are_supports != 0: q is the very point set the grid was built from (conv lists): its cells
// are the work units.  Otherwise the queries are first binned into the support grid's cells inside qws
// (radius_query_ws_bytes(nq, nb)).
int radius_query_cells_dev(const float* q, int64_t nq, const int32_t* q_lens, int64_t ns, int32_t nb, float radius, int32_t width,
                           int32_t row_stride, int32_t* rows, int32_t* counts, int32_t* maxcount, void* ws, size_t ws_bytes, void* qws,
                           size_t qws_bytes, int32_t queries_are_supports, cudaStream_t st)
{
    PCRCG_REQUIRE(nq >= 0 && nq < (1ll << 30), "radius search: Nq out of range");
    Workspace W(ws, ws_bytes);
    RadWS r;
    rad_layout(W, ns, nb, &r);
    PCRCG_REQUIRE(ws != nullptr && W.ok(), "radius search: workspace too small");
    PCRCG_REQUIRE(rows == nullptr || (width >= 0 && row_stride >= width), "radius search: bad row geometry");
    PCRCG_REQUIRE(!queries_are_supports || nq == ns, "radius search: queries_are_supports needs nq == ns");
    if (nq == 0) return PCRCG_OK;
    ProfScope prof(PC_RADIUS_QUERY, st, 8);
    GridWS qg = r.g;
    int koff = 0;
    if (!queries_are_supports) {
        Workspace QW(qws, qws_bytes);
        grid_layout(QW, nq, nb, &qg);
        PCRCG_REQUIRE(qws != nullptr && QW.ok(), "radius search: query workspace too small (%zu < %zu)", qws_bytes, QW.off);
        koff = 2;
        PCRCG_TRY(grid_build(q, (int)nq, q_lens, nb, r.meta, koff, qg, st));
    } else {
        PCRCG_CUDA(cudaMemsetAsync(qg.counters + 1, 0, sizeof(uint32_t), st));      // claim cursor
    }
    if (maxcount) PCRCG_CUDA(cudaMemsetAsync(maxcount, 0, sizeof(int32_t), st));
    const float r2 = radius * radius;      // neighbors.cpp:226 (fp32 product)
    return launch_cells<384, 128, 8>(qg, koff, r, r2, (int)ns, width, row_stride, rows, counts, maxcount, nq, st);
}

// width <= 48 (3DMatch-shaped limits 33-42): cell-centric.  Wider lists mean dense neighbourhoods (the limits are the 80 %
// quantile of the neighbour counts: KITTI-shaped ~100 hits among ~650 candidates): a 1024-candidate / 256-hit configuration of
// the cell kernel fits only 12 warps per SM and measured 6.1 ms per 32-pair KITTI step against 4.5 ms for one warp per query.
bool radius_cells_preferred(int32_t width) { return width > 0 && width <= 48; }

}  // namespace pcrcg
