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
enum ProfClass { PC_SUBSAMPLE = 0, PC_RADIUS_BUILD, PC_RADIUS_QUERY, PC_KPCONV_AGG, PC_GEMM, PC_NORM, PC_POOL, PC_PROJECT, PC_COUNT };
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
#endif

}  // namespace pcrcg
