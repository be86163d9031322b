// mbarrier / bulk-copy (TMA) helpers shared by the kernels that stage data asynchronously
// (sm_90+ PTX; compiled here for sm_100a only).
#pragma once
#include "common.cuh"

#define ASYNC_WAIT_CYCLES (4ll << 30)            // ~2 s: a wait this long is a protocol bug

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u32 bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(u32 bar, u32 tx)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
}
__device__ __forceinline__ bool mbar_try(u32 bar, u32 parity)
{
    u32 ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity)
{
    long long t0 = 0;
    for (u32 spin = 0; !mbar_try(bar, parity); ++spin) {
        if (spin == 0) t0 = clock64();
        else if ((spin & 0x3ffu) == 0 && clock64() - t0 > ASYNC_WAIT_CYCLES) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
