// Common device/host helpers for libpcrcg_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace pcrcg {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define PCRCG_OK 0
#define PCRCG_ERR 1

#define PCRCG_CUDA(expr)                                                                         \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            ::pcrcg::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                \
                               cudaGetErrorString(e__));                                         \
            return PCRCG_ERR;                                                                    \
        }                                                                                        \
    } while (0)

#define PCRCG_TRY(expr)                                                                          \
    do {                                                                                         \
        int r__ = (expr);                                                                        \
        if (r__ != PCRCG_OK) return r__;                                                         \
    } while (0)

#define PCRCG_REQUIRE(cond, ...)                                                                 \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            ::pcrcg::set_error(__VA_ARGS__);                                                     \
            return PCRCG_ERR;                                                                    \
        }                                                                                        \
    } while (0)

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;   // B200

// ---- per-kernel-class device timing (CUDA events on the launching stream) + launch counter ------
// Enabled by pcrcg_profile_enable(1); a ProfScope records an event pair around the launches of one
// kernel class; pcrcg_profile_report() synchronises and sums elapsed times per class.
enum ProfClass { PC_SUBSAMPLE = 0, PC_RADIUS_BUILD, PC_RADIUS_QUERY, PC_KPCONV_AGG, PC_GEMM, PC_NORM, PC_POOL, PC_PROJECT, PC_KPCONV_FUSED, PC_LINEAR, PC_COUNT };
void prof_begin(int cls, cudaStream_t st, int* slot);
void prof_end(int slot, cudaStream_t st);
void count_launches(int n);
struct ProfScope {
    int slot; cudaStream_t st;
    ProfScope(int cls, cudaStream_t s, int launches) : slot(-1), st(s) { count_launches(launches); prof_begin(cls, s, &slot); }
    ~ProfScope() { if (slot >= 0) prof_end(slot, st); }
};

// ---- caller-provided workspace, bump allocated ------------------------------------------------
struct Workspace {
    char* base;
    size_t size;
    size_t off;
    bool dry;   // dry run: only measure
    Workspace(void* p, size_t n) : base((char*)p), size(n), off(0), dry(p == nullptr) {}
    template <class T> T* take(size_t n)
    {
        off = align_up(off, 256);
        T* r = dry ? nullptr : (T*)(base + off);
        off += n * sizeof(T);
        return r;
    }
    bool ok() const { return dry || off <= size; }
};

// Keeps the stream-ordered allocator's memory cached (release threshold = max) -- call before cudaMallocAsync.
int pool_setup();

// ---- primitives (scan_sort.cu) -----------------------------------------------------------------
// out[i] = sum_{j<i} in[j] for i in [0,n]; out has n+1 entries (out[n] = total).  in may alias out.
size_t scan_ws_bytes(int64_t n);
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, void* ws, size_t ws_bytes, cudaStream_t st);

// Stable LSD radix sort of (key,value) u32 pairs on bits [0,nbits).  Result lands in keys_out/vals_out.
size_t sort_ws_bytes(int64_t n);
int radix_sort_pairs(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                     int64_t n, int nbits, void* ws, size_t ws_bytes, cudaStream_t st);

// starts[0..nb] = exclusive scan of lens[0..nb)  (single block)
int cloud_starts(const int32_t* lens, int32_t nb, int32_t* starts, cudaStream_t st);

// ---- device helpers -----------------------------------------------------------------------------
#ifdef __CUDACC__
// index of the cloud containing point i: largest c with starts[c] <= i  (starts has nb+1 entries)
__device__ __forceinline__ int cloud_of(const int32_t* __restrict__ starts, int nb, int i)
{
    int lo = 0, hi = nb;          // invariant: starts[lo] <= i < starts[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(starts + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// Order of the 64 channels of a slab inside the KPConv intermediate when the pipelined aggregation produces it: channel
// c = 8*nt + 2*t + e (accumulator tile nt of mma.m16n8k16, thread column t, element e) is stored at
// 32*(nt>>2) + 8*t + 2*(nt&3) + e, so a thread's four tiles form 16 contiguous bytes and the four threads of a row 64.
// The contraction weights are split with the same permutation of their K index (gemm_tc.cu), so the product is unchanged.
__host__ __device__ __forceinline__ int kperm64(int c)
{
    const int nt = c >> 3, t = (c >> 1) & 3, e = c & 1;
    return ((nt >> 2) << 5) | (t << 3) | ((nt & 3) << 1) | e;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// monotone float <-> int maps for atomicMin/atomicMax on floats
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// Per-cloud bounding boxes (ordered-int encoding, bbox[c] = {min x,y,z, max x,y,z}) of stacked clouds, one point per thread,
// 256-thread blocks.  A warp that lies inside one cloud reduces by shuffle; a block that lies inside one cloud reduces its
// 8 warps through shared memory and issues 6 atomics (instead of 48): the boxes of a stacked batch are a few hundred
// addresses, so the atomics, not the 12 bytes per point, were the cost of this pass.
__device__ __forceinline__ void bbox_accumulate(const float* __restrict__ pts, int n, const int32_t* __restrict__ starts, int nb,
                                                int* __restrict__ bbox)
{
    __shared__ float s_red[8][6];
    __shared__ int s_cloud[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool ok = i < n;
    const int c = ok ? cloud_of(starts, nb, i) : -1;
    float x = 0.f, y = 0.f, z = 0.f;
    if (ok) { x = pts[3 * (size_t)i]; y = pts[3 * (size_t)i + 1]; z = pts[3 * (size_t)i + 2]; }
    const int c0 = __shfl_sync(0xffffffffu, c, 0);
    const bool uniform = __all_sync(0xffffffffu, c == c0);
    float mnx = x, mny = y, mnz = z, mxx = x, mxy = y, mxz = z;
    if (uniform) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o)); mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
        }
    } else if (ok) {                       // a warp that straddles a cloud boundary: per-point atomics (rare)
        int* b = bbox + 6 * c;
        atomicMin(b + 0, f2ord(x)); atomicMin(b + 1, f2ord(y)); atomicMin(b + 2, f2ord(z));
        atomicMax(b + 3, f2ord(x)); atomicMax(b + 4, f2ord(y)); atomicMax(b + 5, f2ord(z));
    }
    if (lane == 0) {
        s_cloud[w] = uniform ? c0 : -2;    // -1: warp past the end, -2: handled above
        s_red[w][0] = mnx; s_red[w][1] = mny; s_red[w][2] = mnz; s_red[w][3] = mxx; s_red[w][4] = mxy; s_red[w][5] = mxz;
    }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        const int cw = lane < nw ? s_cloud[lane] : -1;
        const int cb = __shfl_sync(0xffffffffu, cw, 0);
        // every live warp of the block in the same cloud -> one set of atomics for the block; otherwise one per warp
        const bool block_uniform = cb >= 0 && __all_sync(0xffffffffu, cw == cb || cw == -1);
        if (block_uniform) {
            if (lane < 6) {
                float v = s_red[0][lane];
                for (int k = 1; k < nw; k++)
                    if (s_cloud[k] == cb) v = lane < 3 ? fminf(v, s_red[k][lane]) : fmaxf(v, s_red[k][lane]);
                if (lane < 3) atomicMin(bbox + 6 * cb + lane, f2ord(v)); else atomicMax(bbox + 6 * cb + lane, f2ord(v));
            }
        } else if (lane < nw && cw >= 0) {
            int* b = bbox + 6 * cw;
            atomicMin(b + 0, f2ord(s_red[lane][0])); atomicMin(b + 1, f2ord(s_red[lane][1])); atomicMin(b + 2, f2ord(s_red[lane][2]));
            atomicMax(b + 3, f2ord(s_red[lane][3])); atomicMax(b + 4, f2ord(s_red[lane][4])); atomicMax(b + 5, f2ord(s_red[lane][5]));
        }
    }
}
#endif

}  // namespace pcrcg
