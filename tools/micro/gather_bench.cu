// Micro-benchmark: how fast can a warp stage 32 gathered rows x (128 B hi + 128 B lo) into shared memory?
//   mode 0: cp.async (LDGSTS) 16 B per lane, 16 instructions per 32 rows (what k_kpconv_aggregate_bf16 does)
//   mode 1: cp.async.bulk 128 B, every lane issues its neighbour's two row copies (mbarrier complete_tx)
//   mode 2: cp.async.bulk 128 B, lane 0 issues all 64 copies
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu ; run: ./gather_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>

constexpr int WARPS = 4, ROWS = 32, PITCH = 144;
constexpr int WARP_BYTES = 2 * ROWS * PITCH + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(WARPS * 32) k_gather(const uint8_t* __restrict__ hi, const uint8_t* __restrict__ lo, int ld_bytes,
                                                       const int* __restrict__ idx, int npoints, int H, unsigned* __restrict__ sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint8_t* s_hi = smem + (size_t)w * WARP_BYTES;
    uint8_t* s_lo = s_hi + ROWS * PITCH;
    const uint32_t a_hi = smem_u32(s_hi), a_lo = smem_u32(s_lo);
    const uint32_t bar = smem_u32(s_lo + ROWS * PITCH);
    if (MODE != 0 && lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned acc = 0, phase = 0;
    for (int n = blockIdx.x * WARPS + w; n < npoints; n += gridDim.x * WARPS) {
        const int ja = idx[(size_t)n * H + lane];
        if (MODE == 0) {
            const int chunk = lane & 7, plane = (lane >> 3) & 1, rsel = lane >> 4;
            const uint8_t* xp = (plane ? lo : hi) + chunk * 16;
            const uint32_t dst0 = (plane ? a_lo : a_hi) + (uint32_t)(chunk * 16 + rsel * PITCH);
#pragma unroll
            for (int r = 0; r < ROWS; r += 2) {
                const int j = __shfl_sync(0xffffffffu, ja, r + rsel);
                const void* src = xp + (size_t)j * ld_bytes;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)(r * PITCH)), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        } else {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ROWS * 256) : "memory");
            __syncwarp();
            if (MODE == 1) {
                const uint8_t* sh = hi + (size_t)ja * ld_bytes;
                const uint8_t* sl = lo + (size_t)ja * ld_bytes;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];"
                             ::"r"(a_hi + lane * PITCH), "l"(sh), "r"(bar) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];"
                             ::"r"(a_lo + lane * PITCH), "l"(sl), "r"(bar) : "memory");
            } else {
                for (int r = 0; r < ROWS; r++) {
                    const int j = __shfl_sync(0xffffffffu, ja, r);
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];"
                                     ::"r"(a_hi + r * PITCH), "l"(hi + (size_t)j * ld_bytes), "r"(bar) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];"
                                     ::"r"(a_lo + r * PITCH), "l"(lo + (size_t)j * ld_bytes), "r"(bar) : "memory");
                    }
                }
            }
            asm volatile(
                "{\n\t.reg .pred P1;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                ::"r"(bar), "r"(phase) : "memory");
            phase ^= 1u;
        }
        acc += *reinterpret_cast<const unsigned*>(s_hi + lane * PITCH) + *reinterpret_cast<const unsigned*>(s_lo + lane * PITCH + 64);
        __syncwarp();
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

int main()
{
    const int ns = 40000 * 16, ld = 128, npoints = 646588, H = 32;     // 16 pairs, 64-channel bf16 rows
    uint8_t *hi, *lo; int* idx; unsigned* sink;
    cudaMalloc(&hi, (size_t)ns * ld); cudaMalloc(&lo, (size_t)ns * ld); cudaMalloc(&idx, (size_t)npoints * H * 4); cudaMalloc(&sink, 4);
    cudaMemset(hi, 1, (size_t)ns * ld); cudaMemset(lo, 2, (size_t)ns * ld);
    std::vector<int> h((size_t)npoints * H);
    // neighbours: random rows inside the query's own 40k-row pair (like the hash-ordered clouds)
    uint64_t s = 88172645463325252ull;
    for (int n = 0; n < npoints; n++) {
        int base = (n / 40000) * 40000; if (base + 40000 > ns) base = ns - 40000;
        for (int k = 0; k < H; k++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[(size_t)n * H + k] = base + (int)(s % 40000); }
    }
    cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int smem = WARPS * WARP_BYTES;
    for (int mode = 0; mode < 3; mode++) {
        auto launch = [&](int grid) {
            if (mode == 0) { cudaFuncSetAttribute(k_gather<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_gather<0><<<grid, WARPS * 32, smem>>>(hi, lo, ld, idx, npoints, H, sink); }
            if (mode == 1) { cudaFuncSetAttribute(k_gather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_gather<1><<<grid, WARPS * 32, smem>>>(hi, lo, ld, idx, npoints, H, sink); }
            if (mode == 2) { cudaFuncSetAttribute(k_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_gather<2><<<grid, WARPS * 32, smem>>>(hi, lo, ld, idx, npoints, H, sink); }
        };
        for (int ctas = 2; ctas <= 6; ctas += 1) {
            const int grid = 148 * ctas;
            launch(grid); cudaDeviceSynchronize();
            cudaEventRecord(e0); launch(grid); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            cudaError_t err = cudaGetLastError();
            printf("mode %d  CTAs/SM %d: %.3f ms  %.1f GB/s staged  %.1f ns/point/SM  (%s)\n", mode, ctas, ms, (double)npoints * H * 256 / ms / 1e6,
                   ms * 1e6 / (npoints / 148.0), cudaGetErrorString(err));
        }
    }
    return 0;
}
