// KPConv forward as ONE kernel: neighbour gather -> linear kernel-point influence -> [K*Cin] x Cout contraction on tcgen05,
// with the [Nq, K*Cin] aggregate never leaving the SM.   models/blocks.py:229-374 (rigid kernel, linear influence, sum).
//
// Why this shape.  The aggregate of ONE query point and one 64-channel slab is 15 kernel points x 64 channels = 3.84 KB as
// bf16 (hi, lo) planes; a 128-point UMMA tile of it would be 480 KB, which neither shared memory nor TMEM holds.  So the
// contraction is TRANSPOSED: the weights are the M side and the query points the N side of the MMA,
//
//     D[128, 16] (+)=  A[128, 16]            x  B[16, 16]^T
//                      A = [ W_hi^T ; W_lo^T ]  (64 output channels each, K = 16 of the 960 (kernel point, channel) pairs)
//                      B = [ wf_hi ; wf_lo ]    (8 query points each)
//
// so that one instruction forms all four hi/lo products (the "bf16x3" scheme plus the lo*lo term) and a tile is only 8 query
// points = 30 KB.  The WEIGHTS (64 x 960 x (hi, lo) = 245 KB, more than shared memory) live in TENSOR MEMORY for the whole
// kernel: 128 lanes x 480 columns, written once per CTA with tcgen05.st and read as the TMEM A operand; the remaining 32
// columns are two 16-column accumulators (double buffered).  Nothing but the neighbour rows is re-read per point: no weight
// traffic through shared memory, no intermediate in HBM or L2.
//
// Warp roles (one persistent CTA per SM, 18 warps):
//   warps 5-17  producers: the software-pipelined aggregation of kpconv.cu (k_kpconv_aggregate_bf16p: cp.async double
//               buffering of 16-neighbour k-steps, influence weights in registers, ldmatrix + mma.sync m16n8k16 bf16x3), one
//               query point per warp at a time; the finished [16 kp x 64 ch] fragments are scaled by 1/neighbour count
//               (models/blocks.py:369-372), split hi/lo and stored as row (point % 8) of the tile's 15 x 2 swizzle-128B
//               K-major atoms -- the canonical UMMA operand layout, conflict-free 16-byte stores -- then
//               fence.proxy.async + mbarrier arrive (8 arrivals complete a tile); before a warp tests a slot's "empty" barrier for
//               tile i it waits until tile i - 3 has been issued (tiles_issued, see below): phase-parity waits are only
//               unambiguous one completion ahead
//   warp 4      MMA issuer: per tile 15 x 4 tcgen05.mma.kind::f16 (A from TMEM, B from the tile), tcgen05.commit frees the
//               tile slot and publishes the accumulator
//   warps 0-3   epilogue: tcgen05.ld of the 16 accumulator columns (thread = output channel), hi + lo columns, W_lo rows
//               (lanes 64-127) handed to the W_hi rows through shared memory, coalesced 128-byte row stores, InstanceNorm
//               statistics of the result per (segment, channel) in fp64 registers
//
// Other channel counts (cin = 64 S, cout = 64 T) run as S x T passes of the same kernel (slab s of the features against the
// [s, t] block of the weights, accumulating into the output): correct, but the aggregation is repeated T times, so the
// dispatcher (kpconv.cu) prefers the two-kernel path there -- see DESIGN.md for the measurements.
#include "agg_ptx.cuh"
#include "tc_ptx.cuh"

namespace pcrcg {

constexpr int FZ_PW = 13;                              // producer warps
constexpr int FZ_FIRST_PW = 5;                         // warps 0-3 epilogue, 4 MMA issuer
constexpr int FZ_THREADS = 32 * (FZ_FIRST_PW + FZ_PW);
constexpr int FZ_TILE = 8;                             // query points per tile (N = 16: hi rows | lo rows)
constexpr int FZ_SLOTS = 3;
constexpr int FZ_KP_BYTES = 2048;                      // one kernel point of a tile: hi atom (8 rows x 128 B) + lo atom
constexpr int FZ_SLOT_BYTES = 15 * FZ_KP_BYTES;        // 30 KB
constexpr int FZ_W_COLS = 480;                         // TMEM columns of the weights: 15 kernel points x 32
constexpr int FZ_EPI_BYTES = 2 * FZ_TILE * 64 * 4;     // W_lo partial sums handed between epilogue warps (double buffered)
constexpr int FZ_SMEM = 1024 + FZ_SLOTS * FZ_SLOT_BYTES + FZ_PW * ABP_WARP_BYTES + FZ_EPI_BYTES + 256;

struct FusedStat {
    const int32_t* seg_starts;     // [nseg + 1] absolute row starts (nullptr: no statistics)
    int nseg;
    double* acc;                   // [nseg][2][cout]
    int row0;                      // absolute row of query 0
};

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <typename IdxT>
__global__ void __launch_bounds__(FZ_THREADS, 1) k_kpconv_fused(
    const float* __restrict__ q_pts, int nq, const float* __restrict__ s_pts, int ns, const IdxT* __restrict__ idx, int H, int idx_stride,
    const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo, int cin, int ldxs, const uint8_t* __restrict__ rowflag,
    const float* __restrict__ kpts, int K, float inv_extent, const __nv_bfloat16* __restrict__ w_hi, const __nv_bfloat16* __restrict__ w_lo,
    int ldk, int slab, int accumulate, float* __restrict__ out, int cout, const FusedStat sink)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t stage_base = base + FZ_SLOTS * FZ_SLOT_BYTES;
    const uint32_t epi_base = stage_base + FZ_PW * ABP_WARP_BYTES;
    const uint32_t bar_base = epi_base + FZ_EPI_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (FZ_SLOTS + s); };
    auto accfull_bar = [&](int a) { return bar_base + 8u * (2 * FZ_SLOTS + a); };
    auto accempty_bar = [&](int a) { return bar_base + 8u * (2 * FZ_SLOTS + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * FZ_SLOTS + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));
    int* next_point = reinterpret_cast<int*>(base_ptr + (tmem_slot + 8u - base));      // next unclaimed point of this CTA's sequence
    // Number of tiles whose MMAs have been ISSUED (written by the MMA thread after its commits, read by the producers).  A
    // producer may test the parity of a slot's "empty" barrier for tile i only once tile i - 3 (the slot's previous use) has been
    // issued: then every earlier commit on that barrier has completed (tile i - 3 could only be filled after the commit of tile
    // i - 6), the barrier is at most ONE completion behind and the parity test is unambiguous.  Without this gate a warp that
    // has drifted two uses of a slot ahead (13 warps x 3 claims in flight = 39 points = up to 6 tiles; a 4-k-step point next to
    // 1-k-step points is enough) sees the parity of tile i - 6's completion, overwrites a tile that is still being filled and
    // adds arrivals to its barrier: wrong rows or a hang (seen once in a 4-GPU run of round 2; tests/test_fused_protocol.py
    // reproduces it in a model of these barriers and shows that the gate removes it).
    const uint32_t tiles_issued = tmem_slot + 12u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (nq + FZ_TILE - 1) / FZ_TILE;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int tcol = blockIdx.y;                       // 64-channel block of the output

    if (warp == FZ_FIRST_PW - 1) {
        if (lane == 0) {
            *next_point = 0;
            sts_release(tiles_issued, 0);
            for (int s = 0; s < FZ_SLOTS; s++) { mbar_init(full_bar(s), FZ_TILE); mbar_init(empty_bar(s), 1); }
            for (int a = 0; a < 2; a++) { mbar_init(accfull_bar(a), 1); mbar_init(accempty_bar(a), 4); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    // ---- the weights of this (slab, output block) into tensor memory: lane r < 64 holds W_hi[:, tcol*64 + r], lane 64 + r W_lo;
    //      column 32 kp + c/2 holds K positions (kp, c), (kp, c + 1) of the slab (kperm64 order, as split by gemm_tc.cu) ----
    if (warp < 4) {
        const int r = 32 * warp + lane;
        const __nv_bfloat16* src = ((r >> 6) ? w_lo : w_hi) + (size_t)(tcol * 64 + (r & 63)) * ldk + slab * 64;
        const uint32_t trow = tmem_base + ((uint32_t)(32 * warp) << 16);
        for (int kp = 0; kp < K; kp++) {
            const uint4* p = reinterpret_cast<const uint4*>(src + (size_t)kp * cin);
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
                const uint4 a0 = __ldg(p + 4 * hf), a1 = __ldg(p + 4 * hf + 1), a2 = __ldg(p + 4 * hf + 2), a3 = __ldg(p + 4 * hf + 3);
                const uint32_t v[16] = { a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w };
                tmem_st16(trow + (uint32_t)(kp * 32 + hf * 16), v);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp >= FZ_FIRST_PW) {
        // =================================== producers ===================================
        const int pw = warp - FZ_FIRST_PW;
        const uint32_t a_buf = stage_base + (uint32_t)(pw * ABP_WARP_BYTES);
        float* s_xyz = reinterpret_cast<float*>(base_ptr + (a_buf + 2 * ABP_BUF_BYTES - base));        // [2][16][4]
        const int g = lane >> 2, t = lane & 3;
        const int c0 = slab * 64;
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

        // Points are CLAIMED from a per-CTA counter (sequence number m -> row m % 8 of tile m / 8), three per warp in flight
        // (current, next: indices and query loaded, after next: indices requested).  A static round-robin lets the warps
        // drift apart until the fast ones sit at the edge of the 3-tile ring all the time (measured: 14 % of the producer
        // cycles in the slot wait); claimed in order, the points being finished stay within ~2 tiles of each other.
        const int m_end = my_tiles * FZ_TILE;
        auto claim = [&]() -> int {
            int v = 0;
            if (lane == 0) v = atomicAdd(next_point, 1);
            return __shfl_sync(0xffffffffu, v, 0);
        };
        auto point_of = [&](int m) -> int { return (((int)blockIdx.x + (m >> 3) * (int)gridDim.x) << 3) + (m & 7); };
        auto load_raw = [&](int m, IdxT& r0, IdxT& r1) {
            const int p = m < m_end ? min(point_of(m), nq - 1) : nq - 1;
            const IdxT* row = idx + (size_t)p * idx_stride;
            r0 = row[lane < H ? lane : 0];
            r1 = row[lane + 32 < H ? lane + 32 : 0];
        };
        auto clamp_idx = [&](int m, IdxT r0, IdxT r1, int& j0, int& j1) {
            const bool live = m < m_end && point_of(m) < nq;
            const long long v0 = (long long)r0, v1 = (long long)r1;
            j0 = (live && lane < H && v0 >= 0 && v0 < ns) ? (int)v0 : ns;
            j1 = (live && lane + 32 < H && v1 >= 0 && v1 < ns) ? (int)v1 : ns;
        };
        auto step_mask = [&](int j0, int j1) -> uint32_t {
            const uint32_t m0 = __ballot_sync(0xffffffffu, j0 < ns), m1 = __ballot_sync(0xffffffffu, j1 < ns);
            return ((m0 & 0xffffu) ? 1u : 0u) | ((m0 >> 16) ? 2u : 0u) | ((m1 & 0xffffu) ? 4u : 0u) | ((m1 >> 16) ? 8u : 0u);
        };
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
        auto load_xyz = [&](int j0, int j1, int s, float& x, float& y, float& z) -> bool {
            const int j = __shfl_sync(0xffffffffu, s < 2 ? j0 : j1, ((s & 1) << 4) + (lane & 15));
            const bool v = j < ns;
            if (v && lane < 16) { const float* sp = s_pts + 3 * (size_t)j; x = sp[0]; y = sp[1]; z = sp[2]; }
            return v;
        };
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
        // Row (m % 8) of tile (m / 8): the [16 kp x 64 ch] fragments x 1/count -> bf16 hi / lo -> the tile's swizzle-128B atoms.
        // A thread owns K positions 8t..8t+7 and 32+8t..32+8t+7 of kernel points g and g+8 (kperm64 order): two 16-byte chunks
        // (t and 4+t) per plane and kernel point; odd g stores chunk 4+t first, so the 8 lanes of a store phase hit 8 chunks.
        auto store_point = [&](int m, float inv) {
            const int i = m >> 3, row = m & 7, slot = i % FZ_SLOTS;
            if (i >= FZ_SLOTS && lds_acquire(tiles_issued) < i - (FZ_SLOTS - 1)) {   // the slot's previous tile (i - 3) has been issued
                const long long t0 = clock64();
                uint32_t probes = 0;
                while (lds_acquire(tiles_issued) < i - (FZ_SLOTS - 1)) {
                    __nanosleep(32);
                    if (spin_expired(probes, t0)) __trap();
                }
            }
            mbar_wait(empty_bar(slot), (uint32_t)(((i / FZ_SLOTS) & 1) ^ 1));       // ... and its MMAs are done
            if (point_of(m) < nq) {
                const uint32_t rbase = base + (uint32_t)(slot * FZ_SLOT_BYTES + row * 128);
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    const int kp = g + 8 * hh;
                    uint32_t h[8], l[8];
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) h[nt] = pack_split(acc[nt][2 * hh] * inv, acc[nt][2 * hh + 1] * inv, l[nt]);
                    if (kp < K) {
                        const uint32_t kbase = rbase + (uint32_t)(kp * FZ_KP_BYTES);
                        const uint32_t ca = (uint32_t)(((t + 4 * (g & 1)) ^ row) << 4), cb = (uint32_t)(((t + 4 * ((g & 1) ^ 1)) ^ row) << 4);
                        if (g & 1) {
                            sts128(kbase + ca, h[4], h[5], h[6], h[7]);
                            sts128(kbase + cb, h[0], h[1], h[2], h[3]);
                            sts128(kbase + 1024 + ca, l[4], l[5], l[6], l[7]);
                            sts128(kbase + 1024 + cb, l[0], l[1], l[2], l[3]);
                        } else {
                            sts128(kbase + ca, h[0], h[1], h[2], h[3]);
                            sts128(kbase + cb, h[4], h[5], h[6], h[7]);
                            sts128(kbase + 1024 + ca, l[0], l[1], l[2], l[3]);
                            sts128(kbase + 1024 + cb, l[4], l[5], l[6], l[7]);
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(slot));
#pragma unroll
            for (int i2 = 0; i2 < 8; i2++) { acc[i2][0] = acc[i2][1] = acc[i2][2] = acc[i2][3] = 0.f; }
        };

        int m = claim();
        int nm = claim();
        (void)pw;
        if (m < m_end) {
            int j0, j1;
            {
                IdxT r0, r1;
                load_raw(m, r0, r1);
                clamp_idx(m, r0, r1, j0, j1);
            }
            uint32_t smask = step_mask(j0, j1);
            int s = smask ? __ffs(smask) - 1 : -1;
            float qx, qy, qz;
            {
                const size_t qo = 3 * (size_t)min(point_of(m), nq - 1);
                qx = q_pts[qo]; qy = q_pts[qo + 1]; qz = q_pts[qo + 2];
            }
            int b = 0;
            if (s >= 0) {
                stage(j0, j1, s, 0);
                float x = 0.f, y = 0.f, z = 0.f;
                const bool v = load_xyz(j0, j1, s, x, y, z);
                put_xyz(0, v, x, y, z, qx, qy, qz);
            }
            IdxT nr0, nr1;
            load_raw(nm, nr0, nr1);
            while (m < m_end) {
                const int fm = claim();
                IdxT fr0, fr1;
                load_raw(fm, fr0, fr1);
                int nj0 = ns, nj1 = ns;
                const size_t qo = 3 * (size_t)(nm < m_end ? min(point_of(nm), nq - 1) : 0);
                const float nqx = q_pts[qo], nqy = q_pts[qo + 1], nqz = q_pts[qo + 2];
                const uint8_t f0 = rowflag[j0 < ns ? j0 : 0], f1 = rowflag[j1 < ns ? j1 : 0];
                uint32_t nmask = 0;
                int ns_first = -1;
                if (s < 0) {
                    clamp_idx(nm, nr0, nr1, nj0, nj1);
                    nmask = step_mask(nj0, nj1);
                    ns_first = nmask ? __ffs(nmask) - 1 : -1;
                    if (ns_first >= 0) {
                        stage(nj0, nj1, ns_first, b);
                        float x = 0.f, y = 0.f, z = 0.f;
                        const bool v = load_xyz(nj0, nj1, ns_first, x, y, z);
                        put_xyz(b, v, x, y, z, nqx, nqy, nqz);
                    }
                } else {
                    while (true) {
                        const uint32_t rest = smask & ~((2u << s) - 1u);
                        const bool same = rest != 0u;
                        int s2;
                        if (same) s2 = __ffs(rest) - 1;
                        else {
                            clamp_idx(nm, nr0, nr1, nj0, nj1);
                            nmask = step_mask(nj0, nj1);
                            ns_first = nmask ? __ffs(nmask) - 1 : -1;
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
                    b ^= 1;
                }
                const int cnt = __popc(__ballot_sync(0xffffffffu, j0 < ns && f0 != 0)) + __popc(__ballot_sync(0xffffffffu, j1 < ns && f1 != 0));
                store_point(m, 1.0f / (float)(cnt > 1 ? cnt : 1));
                m = nm;
                nm = fm;
                j0 = nj0; j1 = nj1;
                nr0 = fr0; nr1 = fr1;
                qx = nqx; qy = nqy; qz = nqz;
                smask = nmask;
                s = ns_first;
            }
        }
    } else if (warp == FZ_FIRST_PW - 1) {
        // =================================== MMA issuer ===================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(2 * FZ_TILE);
            for (int i = 0; i < my_tiles; i++) {
                const int slot = i % FZ_SLOTS;
                const uint32_t sph = (uint32_t)((i / FZ_SLOTS) & 1), acc = (uint32_t)(i & 1), aph = (uint32_t)((i >> 1) & 1);
                mbar_wait(accempty_bar(acc), aph ^ 1u);
                mbar_wait(full_bar(slot), sph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + FZ_W_COLS + acc * (2 * FZ_TILE);
                const uint32_t sbase = base + (uint32_t)(slot * FZ_SLOT_BYTES);
                for (int kp = 0; kp < K; kp++) {
                    const uint64_t db = make_desc(sbase + (uint32_t)(kp * FZ_KP_BYTES));
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        umma_bf16_ts(tmem_d, tmem_base + (uint32_t)(kp * 32 + j * 8), db + (uint64_t)((j * 32) >> 4), idesc, (uint32_t)((kp | j) != 0));
                }
                umma_commit(empty_bar(slot));
                umma_commit(accfull_bar(acc));
                sts_release(tiles_issued, i + 1);
            }
        }
    } else {
        // =================================== epilogue ===================================
        const int q = warp;                                          // TMEM lane quarter
        float* s_epi = reinterpret_cast<float*>(base_ptr + (epi_base - base));       // [2][8 points][64 channels]
        const bool stats = sink.acc != nullptr && q < 2;
        const int ch = tcol * 64 + 32 * (q & 1) + lane;              // output channel of this thread (warps 0, 1 and 2, 3 alike)
        double a1 = 0.0, a2 = 0.0;
        int acc_seg = -1, cur_seg = 0;
        auto flush = [&]() {
            if (acc_seg >= 0) {
                atomicAdd(sink.acc + ((size_t)acc_seg * 2) * cout + ch, a1);
                atomicAdd(sink.acc + ((size_t)acc_seg * 2 + 1) * cout + ch, a2);
            }
            a1 = a2 = 0.0;
        };
        for (int i = 0; i < my_tiles; i++) {
            const uint32_t acc = (uint32_t)(i & 1), aph = (uint32_t)((i >> 1) & 1);
            const int n0 = ((int)blockIdx.x + i * (int)gridDim.x) * FZ_TILE;
            mbar_wait(accfull_bar(acc), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t d[16];
            tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + FZ_W_COLS + acc * (2 * FZ_TILE), d);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(accempty_bar(acc));
            float v[FZ_TILE];
#pragma unroll
            for (int p = 0; p < FZ_TILE; p++) v[p] = __uint_as_float(d[p]) + __uint_as_float(d[FZ_TILE + p]);
            float* buf = s_epi + acc * (FZ_TILE * 64);
            if (q >= 2) {
#pragma unroll
                for (int p = 0; p < FZ_TILE; p++) buf[p * 64 + 32 * (q & 1) + lane] = v[p];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (q < 2) {
                const int npts = min(FZ_TILE, nq - n0);
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int p = 0; p < FZ_TILE; p++) {
                    if (p < npts) {
                        float* o = out + (size_t)(n0 + p) * cout + ch;
                        float r = v[p] + buf[p * 64 + 32 * q + lane];
                        if (accumulate) r += *o;
                        *o = r;
                        v[p] = r;
                        s1 += r;
                        s2 = fmaf(r, r, s2);
                    }
                }
                if (stats) {
                    const int first = sink.row0 + n0, last = sink.row0 + n0 + npts - 1;
                    while (cur_seg + 1 < sink.nseg && first >= __ldg(sink.seg_starts + cur_seg + 1)) cur_seg++;
                    const bool uniform = cur_seg + 1 >= sink.nseg || last < __ldg(sink.seg_starts + cur_seg + 1);
                    if (uniform) {
                        if (cur_seg != acc_seg) { flush(); acc_seg = cur_seg; }
                        a1 += (double)s1;
                        a2 += (double)s2;
                    } else {
                        // the 8 points straddle a segment boundary (once per segment): point by point, straight to memory
                        int sg = cur_seg;
#pragma unroll
                        for (int p = 0; p < FZ_TILE; p++) {
                            if (p < npts) {
                                while (sg + 1 < sink.nseg && sink.row0 + n0 + p >= __ldg(sink.seg_starts + sg + 1)) sg++;
                                atomicAdd(sink.acc + ((size_t)sg * 2) * cout + ch, (double)v[p]);
                                atomicAdd(sink.acc + ((size_t)sg * 2 + 1) * cout + ch, (double)v[p] * (double)v[p]);
                            }
                        }
                    }
                }
            }
        }
        if (stats) flush();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == FZ_FIRST_PW - 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

bool kpconv_fused_shape_ok(int64_t nq, int64_t ns, int H, int cin, int cout, int K, int ldxs)
{
    return nq >= 1 && ns >= 1 && H >= 1 && H <= 64 && cin % 64 == 0 && cout % 64 == 0 && K >= 1 && K <= 15 && ldxs >= cin && ldxs % 8 == 0;
}

// w_hi / w_lo: the weights split as for the two-kernel path, bf16 [cout, ldk] with K index kp*cin + (channel, kperm64 inside
// every 64-channel slab).  out [nq, cout] is fully written; statistics (optional) are accumulated into stats_acc.
int kpconv_fused_dev(const float* q_pts, int64_t nq, const float* s_pts, int64_t ns, const void* idx, int idx_is_i64, int H, int idx_stride,
                     const void* x_hi, const void* x_lo, int cin, int ldxs, const uint8_t* rowflag, const float* kpts, int K, float inv_extent,
                     const void* w_hi, const void* w_lo, int ldk, float* out, int cout, const int32_t* seg_starts, int nseg, double* stats_acc,
                     int64_t row0, cudaStream_t st)
{
    PCRCG_REQUIRE(kpconv_fused_shape_ok(nq, ns, H, cin, cout, K, ldxs), "kpconv_fused: unsupported shape");
    PCRCG_REQUIRE(stats_acc == nullptr || (seg_starts != nullptr && nseg >= 1), "kpconv_fused: statistics need segment starts");
    const int S = cin / 64, T = cout / 64;
    const int ntiles = (int)cdiv64(nq, FZ_TILE);
    const dim3 grid((unsigned)(ntiles < kNumSMs ? ntiles : kNumSMs), (unsigned)T);
    if (idx_is_i64) PCRCG_CUDA(cudaFuncSetAttribute(k_kpconv_fused<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM));
    else PCRCG_CUDA(cudaFuncSetAttribute(k_kpconv_fused<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM));
    for (int s = 0; s < S; s++) {
        const FusedStat sink{ seg_starts, nseg, s == S - 1 ? stats_acc : nullptr, (int)row0 };
        count_launches(1);
        if (idx_is_i64)
            k_kpconv_fused<long long><<<grid, FZ_THREADS, FZ_SMEM, st>>>(q_pts, (int)nq, s_pts, (int)ns, (const long long*)idx, H, idx_stride,
                (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, cin, ldxs, rowflag, kpts, K, inv_extent, (const __nv_bfloat16*)w_hi,
                (const __nv_bfloat16*)w_lo, ldk, s, s > 0, out, cout, sink);
        else
            k_kpconv_fused<int><<<grid, FZ_THREADS, FZ_SMEM, st>>>(q_pts, (int)nq, s_pts, (int)ns, (const int*)idx, H, idx_stride,
                (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, cin, ldxs, rowflag, kpts, K, inv_extent, (const __nv_bfloat16*)w_hi,
                (const __nv_bfloat16*)w_lo, ldk, s, s > 0, out, cout, sink);
        PCRCG_CUDA(cudaGetLastError());
    }
    return PCRCG_OK;
}

}  // namespace pcrcg
