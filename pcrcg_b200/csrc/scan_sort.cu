// Device-wide exclusive scan and stable LSD radix sort (u32 key / u32 value), hand written for the
// binning steps of grid subsampling and radius search.  All launches are stream ordered; scratch
// comes from the caller's workspace.
#include "common.cuh"

namespace pcrcg {

// =================================================================================================
// exclusive scan
// =================================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;   // 2048

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Scans one tile per block.  Reads in[i] for i < n_in (0 beyond), writes out[i] for i < n_out.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                             int64_t n_in, int64_t n_out, uint32_t* __restrict__ tile_sums)
{
    __shared__ uint32_t tile[SCAN_TILE];
    __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        int64_t i = base + j * SCAN_THREADS + t;
        tile[j * SCAN_THREADS + t] = i < n_in ? in[i] : 0u;
    }
    __syncthreads();
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) { v[j] = tile[t * SCAN_ITEMS + j]; s += v[j]; }
    uint32_t inc = warp_incl_scan(s, lane);
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; k++) { uint32_t x = warp_tot[k]; if (k < w) woff += x; total += x; }
    uint32_t run = woff + inc - s;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) { tile[t * SCAN_ITEMS + j] = run; run += v[j]; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        int64_t i = base + j * SCAN_THREADS + t;
        if (i < n_out) out[i] = tile[j * SCAN_THREADS + t];
    }
    if (tile_sums != nullptr && t == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t* __restrict__ out, int64_t n_out, const uint32_t* __restrict__ tile_offs)
{
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    const uint32_t off = tile_offs[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        int64_t i = base + j * SCAN_THREADS + threadIdx.x;
        if (i < n_out) out[i] += off;
    }
}

static size_t scan_ws_rec(int64_t n_out)
{
    int64_t tiles = cdiv64(n_out, SCAN_TILE);
    if (tiles <= 1) return 0;
    return align_up((size_t)tiles * sizeof(uint32_t), 256) + scan_ws_rec(tiles);
}

size_t scan_ws_bytes(int64_t n) { return scan_ws_rec(n + 1) + 256; }

static int scan_rec(const uint32_t* in, uint32_t* out, int64_t n_in, int64_t n_out, char* ws, cudaStream_t st)
{
    if (n_out <= 0) return PCRCG_OK;
    int64_t tiles = cdiv64(n_out, SCAN_TILE);
    if (tiles == 1) {
        k_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(in, out, n_in, n_out, nullptr);
        PCRCG_CUDA(cudaGetLastError());
        return PCRCG_OK;
    }
    uint32_t* sums = (uint32_t*)ws;
    char* rest = ws + align_up((size_t)tiles * sizeof(uint32_t), 256);
    k_scan_tiles<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, out, n_in, n_out, sums);
    PCRCG_CUDA(cudaGetLastError());
    PCRCG_TRY(scan_rec(sums, sums, tiles, tiles, rest, st));
    k_scan_add<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(out, n_out, sums);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, void* ws, size_t ws_bytes, cudaStream_t st)
{
    PCRCG_REQUIRE(ws_bytes >= scan_ws_bytes(n) - 256 || n + 1 <= SCAN_TILE, "exclusive_scan_u32: workspace too small");
    return scan_rec(in, out, n, n + 1, (char*)ws, st);
}

// =================================================================================================
// cloud starts (exclusive scan of a short int32 array, single block)
// =================================================================================================
__global__ void __launch_bounds__(1024) k_cloud_starts(const int32_t* __restrict__ lens, int32_t nb, int32_t* __restrict__ starts)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + t;
        uint32_t v = i < nb ? (uint32_t)lens[i] : 0u;
        uint32_t inc = warp_incl_scan(v, lane);
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        uint32_t woff = 0, total = 0;
        for (int k = 0; k < 32; k++) { uint32_t x = warp_tot[k]; if (k < w) woff += x; total += x; }
        uint32_t carry = carry_s;
        if (i < nb) starts[i] = (int32_t)(carry + woff + inc - v);
        __syncthreads();
        if (t == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (t == 0) starts[nb] = (int32_t)carry_s;
}

// out[k] = sum of lens[0 .. k*group) for k in [0, nb/group], and total = sum of all lens (optional): the row starts of
// groups of `group` consecutive clouds (group = 2: fragment pairs) -- single block, nb is small
__global__ void __launch_bounds__(1024) k_group_starts(const int32_t* __restrict__ lens, int nb, int group, int32_t* __restrict__ out,
                                                       int32_t* __restrict__ total)
{
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) { carry_s = 0; out[0] = 0; }
    __syncthreads();
    const int ng = nb / group;
    for (int base = 0; base < ng; base += 1024) {
        const int g = base + t;
        int32_t v = 0;
        if (g < ng) for (int k = 0; k < group; k++) v += lens[g * group + k];
        int32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += x;
        }
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        int32_t woff = 0, tot = 0;
        for (int k = 0; k < 32; k++) { const int32_t x = warp_tot[k]; if (k < w) woff += x; tot += x; }
        const int32_t carry = carry_s;
        if (g < ng) out[g + 1] = carry + woff + inc;
        __syncthreads();
        if (t == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (t == 0 && total != nullptr) {
        int32_t s = carry_s;
        for (int k = ng * group; k < nb; k++) s += lens[k];
        *total = s;
    }
}

int group_starts_dev(const int32_t* lens, int32_t nb, int32_t group, int32_t* out, int32_t* total, cudaStream_t st)
{
    PCRCG_REQUIRE(nb >= 1 && group >= 1 && group <= nb, "group_starts: bad group size");
    count_launches(1);
    k_group_starts<<<1, 1024, 0, st>>>(lens, nb, group, out, total);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

int cloud_starts(const int32_t* lens, int32_t nb, int32_t* starts, cudaStream_t st)
{
    k_cloud_starts<<<1, 1024, 0, st>>>(lens, nb, starts);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

// =================================================================================================
// stable LSD radix sort, 8-bit digits
// =================================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;                      // per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;    // 4096 per block; each warp owns 512 consecutive

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                                        uint32_t* __restrict__ hist, int nblocks)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int j = 0; j < RS_ITEMS; j++) {
        int64_t i = base + j * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                           uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                           int64_t n, int shift, const uint32_t* __restrict__ hist_scanned, int nblocks)
{
    __shared__ uint32_t cnt[RS_WARPS][256];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    for (int k = t; k < RS_WARPS * 256; k += RS_THREADS) (&cnt[0][0])[k] = 0;
    __syncthreads();

    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * (RS_TILE / RS_WARPS);
    uint32_t key[RS_ITEMS], val[RS_ITEMS];
    // phase 1: per-warp digit counts, in element order
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        int64_t i = wbase + j * 32 + lane;
        bool ok = i < n;
        key[j] = ok ? keys_in[i] : 0u;
        val[j] = ok ? vals_in[i] : 0u;
        uint32_t d = ok ? ((key[j] >> shift) & 255u) : 256u;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (ok && (peers & ((1u << lane) - 1u)) == 0) cnt[w][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // phase 2: per digit, exclusive scan over the warps of this block on top of the global base
    {
        uint32_t run = hist_scanned[(size_t)t * nblocks + blockIdx.x];
#pragma unroll
        for (int k = 0; k < RS_WARPS; k++) { uint32_t c = cnt[k][t]; cnt[k][t] = run; run += c; }
    }
    __syncthreads();
    // phase 3: ranked scatter, same order
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        int64_t i = wbase + j * 32 + lane;
        bool ok = i < n;
        uint32_t d = ok ? ((key[j] >> shift) & 255u) : 256u;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t below = peers & ((1u << lane) - 1u);
        if (ok) {
            uint32_t pos = cnt[w][d] + __popc(below);
            keys_out[pos] = key[j];
            vals_out[pos] = val[j];
        }
        __syncwarp();
        if (ok && below == 0) cnt[w][d] += __popc(peers);
        __syncwarp();
    }
}

__global__ void k_copy_pairs(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ oa,
                             uint32_t* __restrict__ ob, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { oa[i] = a[i]; ob[i] = b[i]; }
}

size_t sort_ws_bytes(int64_t n)
{
    int64_t nblocks = cdiv64(n > 0 ? n : 1, RS_TILE);
    size_t hist = align_up((size_t)256 * nblocks * sizeof(uint32_t) + 4, 256);
    return 2 * align_up((size_t)(n > 0 ? n : 1) * sizeof(uint32_t), 256) + hist + scan_ws_bytes(256 * nblocks) + 1024;
}

int radix_sort_pairs(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                     int64_t n, int nbits, void* ws, size_t ws_bytes, cudaStream_t st)
{
    if (n <= 0) return PCRCG_OK;
    PCRCG_REQUIRE(ws_bytes >= sort_ws_bytes(n), "radix_sort_pairs: workspace too small");
    Workspace W(ws, ws_bytes);
    int nblocks = (int)cdiv64(n, RS_TILE);
    uint32_t* tk = W.take<uint32_t>(n);
    uint32_t* tv = W.take<uint32_t>(n);
    uint32_t* hist = W.take<uint32_t>((size_t)256 * nblocks + 1);
    W.off = align_up(W.off, 256);
    void* sws = W.base + W.off;
    size_t sws_bytes = W.size - W.off;

    int passes = (nbits + 7) / 8;
    if (passes == 0) {
        k_copy_pairs<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(keys_in, vals_in, keys_out, vals_out, n);
        PCRCG_CUDA(cudaGetLastError());
        return PCRCG_OK;
    }
    const uint32_t* ck = keys_in;
    const uint32_t* cv = vals_in;
    for (int p = 0; p < passes; p++) {
        bool to_out = ((passes - 1 - p) % 2) == 0;
        uint32_t* ok_ = to_out ? keys_out : tk;
        uint32_t* ov_ = to_out ? vals_out : tv;
        k_rs_hist<<<nblocks, RS_THREADS, 0, st>>>(ck, n, p * 8, hist, nblocks);
        PCRCG_CUDA(cudaGetLastError());
        PCRCG_TRY(exclusive_scan_u32(hist, hist, (int64_t)256 * nblocks, sws, sws_bytes, st));
        k_rs_scatter<<<nblocks, RS_THREADS, 0, st>>>(ck, cv, ok_, ov_, n, p * 8, hist, nblocks);
        PCRCG_CUDA(cudaGetLastError());
        ck = ok_;
        cv = ov_;
    }
    return PCRCG_OK;
}

}  // namespace pcrcg
