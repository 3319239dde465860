// Warp-level helpers of the bf16x3 KPConv aggregation (mma.sync m16n8k16 + ldmatrix), shared by kpconv.cu and kpconv_fused.cu.
#pragma once
#include "common.cuh"

#include <cuda_bf16.h>

namespace pcrcg {

constexpr int KP_MAX = 16;                     // kernel points padded to 16 (the M of mma.sync)
constexpr int AB_PITCH = 144;                  // bytes per staged plane row (128 + 16: conflict-free ldmatrix)
constexpr int ABP_ROWS = 16;                   // neighbours per k-step of the pipelined kernels
constexpr int ABP_BUF_BYTES = 2 * ABP_ROWS * AB_PITCH;                        // hi + lo rows of one k-step
constexpr int ABP_WARP_BYTES = 2 * ABP_BUF_BYTES + 2 * ABP_ROWS * 16;         // two buffers + two coordinate blocks

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void stmatrix_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3)
{
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
// Splits two fp32 values into packed bf16 (hi, lo) pairs (low half = a, high half = b) WITHOUT conversion instructions:
// hi = the upper 16 bits of the float (truncation), lo = the upper 16 bits of the exact residual x - hi.  hi + lo keeps
// >= 15 mantissa bits (error <= 2^-16 |x|, of the same order as the lo*lo term bf16x3 drops); PRMT / LOP3 / FADD run on the
// full-rate pipes where F2F (one per value and rounding step, 112 per point before) is quarter-rate.
__device__ __forceinline__ uint32_t pack_split(float a, float b, uint32_t& lo_packed)
{
    const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
    const float ra = a - __uint_as_float(ua & 0xffff0000u), rb = b - __uint_as_float(ub & 0xffff0000u);
    lo_packed = __byte_perm(__float_as_uint(ra), __float_as_uint(rb), 0x7632);
    return __byte_perm(ua, ub, 0x7632);
}

}  // namespace pcrcg
