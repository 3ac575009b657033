// b2_common.cuh -- device/host helpers shared by the sm_100a kernels of the encode stage.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/b2enc_types.h"

#define B2_CUDA_OK(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            fprintf(stderr, "b2enc: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),  \
                    __FILE__, __LINE__, cudaGetErrorString(_e));                          \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)

// ---- host: TMA descriptor over a stack of padded u8 planes [n][rows][pitch] -------------
// Returns 0 on success.  box = {bw, bh, 1}.
int b2_make_plane_tmap(CUtensorMap *tm, const void *base, int pitch, int rows, int nplanes,
                       int bw, int bh);

#ifdef __CUDACC__
// ---- device: byte SIMD -------------------------------------------------------------------
// one VABSDIFF4.U8.ACC: acc + sum_{b<4} |a.b - b.b|
__device__ __forceinline__ uint32_t vsad4_acc(uint32_t a, uint32_t b, uint32_t acc)
{
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}

__device__ __forceinline__ int b2_mvbits(int v)
{
    if (v == 0) return 1;
    return 2 * (31 - __clz(abs(v))) + 3;
}

__device__ __forceinline__ int b2_clip255(int v) { return min(max(v, 0), 255); }

// ---- device: mbarrier + TMA --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "B2_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra B2_DONE_%=;\n"
        "bra B2_WAIT_%=;\n"
        "B2_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// 3-D tiled TMA load: box at element coordinates (c0,c1,c2) -> dense smem tile
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
#endif
