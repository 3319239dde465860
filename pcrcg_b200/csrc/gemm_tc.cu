// tcgen05 tensor-core contraction for the KPConv weight contraction and the unary Linear layers.
//
//   C[M,N] = A[M,K] * B^T  (B stored [N,K], K contiguous), fp32 in / fp32 out, row scale epilogue.
//
// Precision: operands are split on the fly into bf16 (hi, lo) pairs and the product is evaluated as
//   A_hi*B_hi + A_hi*B_lo + A_lo*B_hi      ("bf16x3", fp32 accumulation in TMEM)
// which keeps ~16 mantissa bits per operand (measured ~1e-5 normwise vs fp32), inside the 1e-3
// feature tolerance with margin where a single bf16 (2e-3) or tf32 (3e-4 per op, 11 stacked blocks)
// pass is not.
//
// Kernel (PERSISTENT: one CTA per SM loops over 128 x BN output tiles, 192 or 320 threads; the accumulator is
// double buffered in TMEM so the epilogue of tile i overlaps the TMA loads and MMAs of tile i+1):
//   warp 0   : TMA producer -- cp.async.bulk.tensor 2D loads of the four operand tiles of a
//              64-wide K block (A_hi, A_lo [128 x 64], B_hi, B_lo [BN x 64], 128B swizzle) into a
//              STAGES-deep shared-memory ring, mbarrier expect_tx / complete_tx
//   warp 1   : MMA issuer -- one elected lane issues tcgen05.mma.kind::f16 (M=128, K=16):
//                 D[:, 0:2BN] += A_hi * [B_hi ; B_lo]^T      (one instruction, N = 2*BN)
//                 D[:, 0:BN]  += A_lo * B_hi^T
//              accumulators live in TMEM (2*BN fp32 columns); tcgen05.commit frees the smem slot
//   warps 2-5 (2-9 for BN >= 64): epilogue -- tcgen05.ld 32x32b, (D1 + D2) * row_scale, shared-memory transpose, row-segment
//              stores to global; optionally the per-(segment, column) sum / sum of squares of the tile for the InstanceNorm
//              that follows every contraction of the path (saves a full read pass over C)
#include "tc_ptx.cuh"

namespace pcrcg {

template <int BN> struct TcCfg {
    static constexpr int A_BYTES = TC_BM * TC_BK * 2;               // 16 KB per (hi|lo)
    static constexpr int B_BYTES = BN * TC_BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
    static constexpr int ACC_COLS = 2 * BN;                          // D1 | D2 of one accumulator set
    static constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;   // two sets: 64, 128, 256 or 512 columns
    static constexpr int EPI_WARPS = BN >= 64 ? 8 : 4;               // two warps per TMEM lane quarter for wide tiles
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int STAT_BYTES = EPI_WARPS * 32 * 20 * 4;       // per epilogue warp: 32 rows x 16 columns (pitch 20 floats) staging
    static constexpr int SEG_CACHE = 256;                            // segment starts cached in shared memory (else read from global)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + STAT_BYTES + (SEG_CACHE + 4) * 4;
};

// Column statistics sink of the epilogue: per (segment, column) sum and sum of squares of the values written,
// added in fp64 to acc[nseg][2][N] (InstanceNorm statistics without a second pass over C; dense.cu finalises them).
struct StatSink {
    const int32_t* seg_starts;     // [nseg + 1] absolute row starts
    int nseg;
    double* acc;                   // nullptr = no statistics
    int row0;                      // absolute row of C's row 0 (chunked callers)
    int dbg;                       // measurement switch: 1 = no atomics, 2 = no column sums
};
static int g_stats_dbg = 0;
void gemm_set_stats_dbg(int v) { g_stats_dbg = v; }

template <int BN>
__global__ void __launch_bounds__(TcCfg<BN>::THREADS, 1) k_gemm_bf16x3(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                                                             const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                                                             float* __restrict__ C, int ldc, int M, int N, int K,
                                                             const float* __restrict__ row_scale, int n_tiles_n, int total_tiles,
                                                             const StatSink sink)
{
    using Cfg = TcCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + Cfg::STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (K + TC_BK - 1) / TC_BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo));
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), Cfg::EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // segment starts of the statistics sink: a shared-memory copy keeps the per-tile segment lookup of the epilogue off the
    // (heavily loaded) global-memory path
    int32_t* seg_cache = reinterpret_cast<int32_t*>(smem_raw + (bar_base + 256u + (uint32_t)Cfg::STAT_BYTES - smem_u32(smem_raw)));
    const bool seg_cached = sink.acc != nullptr && sink.nseg <= Cfg::SEG_CACHE;
    if (seg_cached)
        for (int i = threadIdx.x; i <= sink.nseg; i += blockDim.x) seg_cache[i] = sink.seg_starts[i];
    const int32_t* segp = seg_cached ? seg_cache : sink.seg_starts;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles_n) * TC_BM, n0 = (tile % n_tiles_n) * BN;
                for (int kb = 0; kb < num_kb; kb++, it++) {
                    const int s = it % Cfg::STAGES;
                    const uint32_t ph = (it / Cfg::STAGES) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t st = base + s * Cfg::STAGE_BYTES;
                    mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
                    tma_load_2d(st, &map_a_hi, full_bar(s), kb * TC_BK, m0);
                    tma_load_2d(st + Cfg::A_BYTES, &map_a_lo, full_bar(s), kb * TC_BK, m0);
                    tma_load_2d(st + 2 * Cfg::A_BYTES, &map_b_hi, full_bar(s), kb * TC_BK, n0);
                    tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &map_b_lo, full_bar(s), kb * TC_BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc_cat = make_idesc(2 * BN);
            constexpr uint32_t idesc_one = make_idesc(BN);
            uint32_t it = 0, t = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, t++) {
                const uint32_t acc = t & 1u, aph = (t >> 1) & 1u;
                mbar_wait(tmem_empty_bar(acc), aph ^ 1u);          // epilogue has drained this accumulator set
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_COLS;
                for (int kb = 0; kb < num_kb; kb++, it++) {
                    const int s = it % Cfg::STAGES;
                    const uint32_t ph = (it / Cfg::STAGES) & 1u;
                    mbar_wait(full_bar(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * Cfg::STAGE_BYTES;
                    const uint64_t da_hi = make_desc(st), da_lo = make_desc(st + Cfg::A_BYTES), db = make_desc(st + 2 * Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; k++) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);      // 16 bf16 = 32 bytes inside the swizzle atom
                        umma_bf16(tmem_d, da_hi + adv, db + adv, idesc_cat, (uint32_t)((kb | k) != 0));
                        umma_bf16(tmem_d, da_lo + adv, db + adv, idesc_one, 1u);
                    }
                    umma_commit(empty_bar(s));
                }
                umma_commit(tmem_full_bar(acc));
            }
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), +32); with 8 epilogue warps (BN >= 64) the two warps of a
        // lane quarter split the tile's columns =====
        // Per chunk of 16 columns: tcgen05.ld (the next chunk's loads are in flight while this one is processed) -> (D1 + D2)
        // * row_scale -> the warp's 32 x 16 block is staged in shared memory so that (a) global stores are whole row segments
        // (8 rows x 64 B per instruction instead of 32 scattered 16-byte pieces) and (b) columns can be summed for the
        // InstanceNorm statistics.  Statistics stay in per-lane fp64 registers while the CTA remains inside one segment
        // (its n-tile never changes: the grid is a multiple of n_tiles_n) and are flushed with one atomic per column.
        constexpr int EW = Cfg::EPI_WARPS;
        constexpr int WCOLS = EW == 8 ? BN / 2 : BN;               // columns of the tile handled by one warp
        constexpr int CH = 16, NCH = WCOLS / CH;
        constexpr int PITCH = 20;                                  // floats: 16-byte aligned rows; conflict-free STS.128 / LDS.128
        const int ew = warp - 2, q = warp & 3;
        const int wc0 = EW == 8 ? (ew >> 2) * WCOLS : 0;           // first column (inside the tile) of this warp
        float* stage = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw))) + ew * (32 * PITCH);
        const bool stats = sink.acc != nullptr;
        double a1[NCH], a2[NCH];                                   // lanes 0-15: column sums, lanes 16-31: sums of squares
#pragma unroll
        for (int i = 0; i < NCH; i++) a1[i] = 0.0;
        (void)a2;
        int acc_seg = -1, acc_n0 = 0, cur_seg = 0;
        auto flush = [&]() {
            if (acc_seg >= 0 && !(sink.dbg & 1)) {
#pragma unroll
                for (int i = 0; i < NCH; i++) {
                    const int col = acc_n0 + wc0 + i * CH + (lane & 15);
                    if (col < N) atomicAdd(sink.acc + ((size_t)acc_seg * 2 + (lane >> 4)) * N + col, a1[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < NCH; i++) a1[i] = 0.0;
        };
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, t++) {
            const int m0 = (tile / n_tiles_n) * TC_BM, n0 = (tile % n_tiles_n) * BN;
            const uint32_t acc = t & 1u, aph = (t >> 1) & 1u;
            const int wrow0 = m0 + q * 32, row = wrow0 + lane;
            const bool wvalid = wrow0 < M;
            int seg = 0;
            bool seg_uniform = true;
            if (stats && wvalid) {
                if (sink.nseg > 1) {
                    // m0 only grows: advance the segment cursor (warp-uniform, normally zero or one step)
                    const int first = sink.row0 + wrow0, last = sink.row0 + min(wrow0 + 31, M - 1);
                    while (cur_seg + 1 < sink.nseg && first >= segp[cur_seg + 1]) cur_seg++;
                    seg = cur_seg;
                    seg_uniform = cur_seg + 1 >= sink.nseg || last < segp[cur_seg + 1];
                    if (!seg_uniform) {               // rare: per-row lookup (generic loads: segp may point to shared memory)
                        const int r = sink.row0 + min(row, M - 1);
                        seg = cur_seg;
                        while (seg + 1 < sink.nseg && r >= segp[seg + 1]) seg++;
                    }
                }
                if (seg_uniform && (seg != acc_seg || n0 != acc_n0)) {
                    flush();
                    acc_seg = seg;
                    acc_n0 = n0;
                }
            }
            mbar_wait(tmem_full_bar(acc), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const float sc = (row_scale != nullptr && row < M) ? row_scale[row] : 1.0f;
            const uint32_t trow = tmem_base + acc * Cfg::ACC_COLS + ((uint32_t)(q * 32) << 16) + (uint32_t)wc0;
            uint32_t d[2][2 * CH];
            tmem_ld16(trow, d[0]);
            tmem_ld16(trow + (uint32_t)BN, d[0] + CH);
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float v[CH];
#pragma unroll
                for (int u = 0; u < CH; u++) v[u] = __uint_as_float(d[ch & 1][u]) + __uint_as_float(d[ch & 1][CH + u]);
                if (row_scale != nullptr) {                   // warp-uniform: the Linears (no row scale) skip the multiplies
#pragma unroll
                    for (int u = 0; u < CH; u++) v[u] *= sc;
                }
                if (ch + 1 < NCH) {
                    tmem_ld16(trow + (uint32_t)((ch + 1) * CH), d[(ch + 1) & 1]);
                    tmem_ld16(trow + (uint32_t)(BN + (ch + 1) * CH), d[(ch + 1) & 1] + CH);
                }
#pragma unroll
                for (int u = 0; u < CH; u += 4)
                    *reinterpret_cast<float4*>(stage + lane * PITCH + u) = make_float4(v[u], v[u + 1], v[u + 2], v[u + 3]);
                __syncwarp();
                const int cbase = n0 + wc0 + ch * CH;
                if ((ldc & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {                       // 8 rows x 64 bytes per instruction
                        const int r = j * 8 + (lane >> 2), piece = (lane & 3) * 4;
                        if (wrow0 + r < M && cbase + piece < N)
                            *reinterpret_cast<float4*>(C + (size_t)(wrow0 + r) * ldc + cbase + piece) =
                                *reinterpret_cast<const float4*>(stage + r * PITCH + piece);
                    }
                } else if (lane < CH && cbase + lane < N) {
                    for (int r = 0; r < 32 && wrow0 + r < M; r++) C[(size_t)(wrow0 + r) * ldc + cbase + lane] = stage[r * PITCH + lane];
                }
                if (stats && wvalid && !(sink.dbg & 2)) {
                    const int cl = lane & 15, rh = lane >> 4;
                    if (seg_uniform) {
                        // half-warp rh sums rows (i/4)*8 + i%4 + 4*rh (bank-conflict free with pitch 20); rows >= M hold exact
                        // zeros (TMA zero fill of A)
                        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const float x = stage[((i >> 2) * 8 + (i & 3) + 4 * rh) * PITCH + cl];
                            const float y = stage[(((i + 1) >> 2) * 8 + ((i + 1) & 3) + 4 * rh) * PITCH + cl];
                            s1a += x; s2a = fmaf(x, x, s2a);
                            s1b += y; s2b = fmaf(y, y, s2b);
                        }
                        float s1 = s1a + s1b, s2 = s2a + s2b;
                        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
                        a1[ch] += (double)(rh ? s2 : s1);
                    } else {
                        // the warp's rows straddle a segment boundary: one masked column sum per segment, straight to memory
                        const int sfirst = __shfl_sync(0xffffffffu, seg, 0), slast = __shfl_sync(0xffffffffu, seg, 31);
                        for (int sg = sfirst; sg <= slast; sg++) {
                            float s1 = 0.f, s2 = 0.f;
                            for (int i = 0; i < 32; i++) {
                                const int si = __shfl_sync(0xffffffffu, seg, i);
                                const float x = stage[i * PITCH + cl];
                                if (si == sg && wrow0 + i < M) { s1 += x; s2 = fmaf(x, x, s2); }
                            }
                            if (cbase + cl < N && !(sink.dbg & 1))
                                atomicAdd(sink.acc + ((size_t)sg * 2 + rh) * N + cbase + cl, (double)(rh ? s2 : s1));
                        }
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        }
        if (stats) flush();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// operand splitting: x = hi + lo with hi = bf16(x), lo = bf16(x - hi)
__global__ void __launch_bounds__(256) k_split_bf16(const float* __restrict__ x, int ldx, int rows, int cols, __nv_bfloat16* __restrict__ hi,
                                                    __nv_bfloat16* __restrict__ lo, int ldo)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4 = ldo >> 2;
    if (e >= (long long)rows * c4) return;
    const int r = (int)(e / c4), c = (int)(e - (long long)r * c4) * 4;
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = (c + u < cols) ? x[(size_t)r * ldx + c + u] : 0.f;
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        h[u] = __float2bfloat16_rn(v[u]);
        l[u] = __float2bfloat16_rn(v[u] - __bfloat162float(h[u]));
    }
    *reinterpret_cast<uint2*>(hi + (size_t)r * ldo + c) = *reinterpret_cast<uint2*>(h);
    *reinterpret_cast<uint2*>(lo + (size_t)r * ldo + c) = *reinterpret_cast<uint2*>(l);
}

// B given as [K,N] row-major -> hi/lo [N, ldo] (K contiguous), via a 32x32 shared-memory transpose.
// perm64: the K index is permuted inside every aligned block of 64 (common.cuh: kperm64) -- the order in which the
// pipelined KPConv aggregation writes its 64-channel slabs (one thread's accumulator fragments become contiguous).
__global__ void __launch_bounds__(256) k_split_bf16_transpose(const float* __restrict__ B, int ldb, int K, int N, __nv_bfloat16* __restrict__ hi,
                                                              __nv_bfloat16* __restrict__ lo, int ldo, int perm64)
{
    __shared__ float t[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        int k = k0 + j, n = n0 + tx;
        t[j][tx] = (k < K && n < N) ? B[(size_t)k * ldb + n] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int n = n0 + j, k = k0 + tx;
        if (n < N && k < ldo) {
            float v = k < K ? t[tx][j] : 0.f;
            const int kk = perm64 ? ((k & ~63) | kperm64(k & 63)) : k;
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            hi[(size_t)n * ldo + kk] = h;
            lo[(size_t)n * ldo + kk] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int get_encode()
{
    if (g_encode) return PCRCG_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PCRCG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    PCRCG_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is unavailable in this driver");
    g_encode = (EncodeTiledFn)fn;
    return PCRCG_OK;
}

// 2D bf16 tensor [rows, cols] (cols contiguous, pitch ld elements), box [box_rows x 64], 128B swizzle
static int make_map(CUtensorMap* m, const void* ptr, int rows, int cols, int ld, int box_rows)
{
    cuuint64_t dims[2] = { (cuuint64_t)cols, (cuuint64_t)rows };
    cuuint64_t strides[1] = { (cuuint64_t)ld * 2 };
    cuuint32_t box[2] = { (cuuint32_t)TC_BK, (cuuint32_t)box_rows };
    cuuint32_t estr[2] = { 1, 1 };
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PCRCG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld);
    return PCRCG_OK;
}

template <int BN>
static int launch_tc(const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, const __nv_bfloat16* b_hi, const __nv_bfloat16* b_lo, int ldk,
                     float* C, int ldc, int M, int N, int K, const float* row_scale, cudaStream_t st, const StatSink& sink)
{
    using Cfg = TcCfg<BN>;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    PCRCG_TRY(make_map(&ma_hi, a_hi, M, K, ldk, TC_BM));
    PCRCG_TRY(make_map(&ma_lo, a_lo, M, K, ldk, TC_BM));
    PCRCG_TRY(make_map(&mb_hi, b_hi, N, K, ldk, BN));
    PCRCG_TRY(make_map(&mb_lo, b_lo, N, K, ldk, BN));
    // per device (function attributes live in the context): cheap enough to set on every launch
    PCRCG_CUDA(cudaFuncSetAttribute(k_gemm_bf16x3<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    const int n_tiles_n = (int)cdiv64(N, BN);
    const long long total = cdiv64(M, TC_BM) * n_tiles_n;
    PCRCG_REQUIRE(total < (1ll << 31), "gemm_tc: too many tiles");
    // persistent: one CTA per SM; a multiple of n_tiles_n so that a CTA keeps its n-tile (epilogue statistics stay in registers)
    unsigned grid = (unsigned)(total < kNumSMs ? total : kNumSMs);
    if ((int)grid > n_tiles_n) grid -= grid % (unsigned)n_tiles_n;
    k_gemm_bf16x3<BN><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, C, ldc, M, N, K, row_scale, n_tiles_n, (int)total, sink);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

static bool g_pool_ready[64] = { false };       // per device

int pool_setup()
{
    int dev = 0;
    PCRCG_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && g_pool_ready[dev]) return PCRCG_OK;
    cudaMemPool_t pool;
    PCRCG_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t thr = ~0ull;
    PCRCG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    if (dev >= 0 && dev < 64) g_pool_ready[dev] = true;
    return PCRCG_OK;
}

bool gemm_tc_shape_ok(int M, int N, int K) { return N % 16 == 0 && N >= 16 && K >= 16 && M >= 1; }

int split_bf16_dev(const float* x, int ldx, int64_t rows, int cols, void* hi, void* lo, int ldo, cudaStream_t st)
{
    PCRCG_REQUIRE(ldo % 8 == 0 && ldo >= cols && rows >= 0, "split_bf16: bad geometry");
    if (rows == 0) return PCRCG_OK;
    count_launches(1);
    long long tot = (long long)rows * (ldo / 4);
    k_split_bf16<<<(unsigned)cdiv64(tot, 256), 256, 0, st>>>(x, ldx, (int)rows, cols, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ldo);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

// Split (and transpose if needed) the fp32 B operand into bf16 hi/lo [N, ldk] (K contiguous).
int gemm_tc_split_b_perm_dev(const float* B, int ldb, int b_is_nk, int N, int K, int ldk, void* b_hi_v, void* b_lo_v, int perm64, cudaStream_t st);
int gemm_tc_split_b_dev(const float* B, int ldb, int b_is_nk, int N, int K, int ldk, void* b_hi_v, void* b_lo_v, cudaStream_t st)
{
    return gemm_tc_split_b_perm_dev(B, ldb, b_is_nk, N, K, ldk, b_hi_v, b_lo_v, 0, st);
}

int gemm_tc_split_b_perm_dev(const float* B, int ldb, int b_is_nk, int N, int K, int ldk, void* b_hi_v, void* b_lo_v, int perm64, cudaStream_t st)
{
    PCRCG_REQUIRE(!perm64 || (!b_is_nk && K % 64 == 0), "gemm_tc: the permuted split needs a [K,N] operand with K %% 64 == 0");
    __nv_bfloat16* b_hi = (__nv_bfloat16*)b_hi_v;
    __nv_bfloat16* b_lo = (__nv_bfloat16*)b_lo_v;
    count_launches(1);
    if (b_is_nk) {
        long long tot = (long long)N * (ldk / 4);
        k_split_bf16<<<(unsigned)cdiv64(tot, 256), 256, 0, st>>>(B, ldb, N, K, b_hi, b_lo, ldk);
    } else {
        k_split_bf16_transpose<<<dim3((unsigned)cdiv64(ldk, 32), (unsigned)cdiv64(N, 32)), 256, 0, st>>>(B, ldb, K, N, b_hi, b_lo, ldk, perm64);
    }
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

// Both operands already split: a_* bf16 [M, ldk], b_* bf16 [N, ldk] (ldk % 8 == 0).
int gemm_tc_core_stats_dev(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int ldk, float* C, int ldc, int M, int N,
                           int K, const float* row_scale, cudaStream_t st, const int32_t* seg_starts, int nseg, double* stats_acc, int64_t row0)
{
    StatSink sink{ seg_starts, nseg, stats_acc, (int)row0, g_stats_dbg };
    PCRCG_REQUIRE(stats_acc == nullptr || (seg_starts != nullptr && nseg >= 1), "gemm_tc: statistics need segment starts");
    PCRCG_REQUIRE(gemm_tc_shape_ok(M, N, K) && ldk % 8 == 0 && ldk >= K, "gemm_tc: unsupported shape M=%d N=%d K=%d ldk=%d", M, N, K, ldk);
    PCRCG_TRY(get_encode());
    count_launches(1);
    const __nv_bfloat16 *ah = (const __nv_bfloat16*)a_hi, *al = (const __nv_bfloat16*)a_lo, *bh = (const __nv_bfloat16*)b_hi,
                        *bl = (const __nv_bfloat16*)b_lo;
    if (N % 128 == 0) return launch_tc<128>(ah, al, bh, bl, ldk, C, ldc, M, N, K, row_scale, st, sink);
    if (N % 64 == 0) return launch_tc<64>(ah, al, bh, bl, ldk, C, ldc, M, N, K, row_scale, st, sink);
    if (N % 32 == 0) return launch_tc<32>(ah, al, bh, bl, ldk, C, ldc, M, N, K, row_scale, st, sink);
    return launch_tc<16>(ah, al, bh, bl, ldk, C, ldc, M, N, K, row_scale, st, sink);
}

int gemm_tc_core_dev(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int ldk, float* C, int ldc, int M, int N, int K,
                     const float* row_scale, cudaStream_t st)
{
    return gemm_tc_core_stats_dev(a_hi, a_lo, b_hi, b_lo, ldk, C, ldc, M, N, K, row_scale, st, nullptr, 0, nullptr, 0);
}

// A already split: a_hi / a_lo bf16 [M, ldk] (ldk % 8 == 0).  B fp32, split here (weights are small).
int gemm_tc_presplit_dev(const void* a_hi_v, const void* a_lo_v, int ldk, const float* B, int ldb, int b_is_nk, float* C, int ldc, int M,
                         int N, int K, const float* row_scale, cudaStream_t st)
{
    PCRCG_REQUIRE(gemm_tc_shape_ok(M, N, K) && ldk % 8 == 0 && ldk >= K, "gemm_tc: unsupported shape M=%d N=%d K=%d ldk=%d", M, N, K, ldk);
    PCRCG_TRY(pool_setup());
    __nv_bfloat16* b_hi = nullptr;
    const size_t b_elems = (size_t)N * ldk;
    PCRCG_CUDA(cudaMallocAsync((void**)&b_hi, 2 * b_elems * sizeof(__nv_bfloat16) + 1024, st));
    __nv_bfloat16* b_lo = b_hi + b_elems;
    int rc = gemm_tc_split_b_dev(B, ldb, b_is_nk, N, K, ldk, b_hi, b_lo, st);
    if (rc == PCRCG_OK) rc = gemm_tc_core_dev(a_hi_v, a_lo_v, b_hi, b_lo, ldk, C, ldc, M, N, K, row_scale, st);
    cudaFreeAsync(b_hi, st);
    return rc;
}

int gemm_tc_dev(const float* A, int lda, const float* B, int ldb, int b_is_nk, float* C, int ldc, int M, int N, int K,
                const float* row_scale, cudaStream_t st, bool* handled)
{
    *handled = false;
    if (!gemm_tc_shape_ok(M, N, K)) return PCRCG_OK;        // ragged shapes stay on the CUDA-core kernel
    PCRCG_TRY(pool_setup());
    const int ldk = (K + 7) / 8 * 8;                       // bf16 row pitch: multiple of 16 bytes for TMA
    __nv_bfloat16* a_hi = nullptr;
    const size_t a_elems = (size_t)M * ldk;
    PCRCG_CUDA(cudaMallocAsync((void**)&a_hi, 2 * a_elems * sizeof(__nv_bfloat16) + 1024, st));
    __nv_bfloat16* a_lo = a_hi + a_elems;
    count_launches(1);
    long long tot = (long long)M * (ldk / 4);
    k_split_bf16<<<(unsigned)cdiv64(tot, 256), 256, 0, st>>>(A, lda, M, K, a_hi, a_lo, ldk);
    int rc = PCRCG_OK;
    if (cudaGetLastError() != cudaSuccess) { set_error("gemm_tc: split launch failed"); rc = PCRCG_ERR; }
    if (rc == PCRCG_OK) rc = gemm_tc_presplit_dev(a_hi, a_lo, ldk, B, ldb, b_is_nk, C, ldc, M, N, K, row_scale, st);
    cudaFreeAsync(a_hi, st);
    if (rc == PCRCG_OK) *handled = true;
    return rc;
}

}  // namespace pcrcg
