// k9_pack.cu -- K9: compaction of the quantised levels before they leave the GPU (row a9 of SURVEY.md 8a, the
// hand-off to the serial entropy stage).
//
// The dense record is 832 B per macroblock (26 blocks x 16 int16), but at the QPs of BASELINE.json's configs a P
// macroblock holds only ~3.5 blocks with a non-zero level: shipping the dense array makes the device->host copy
// (7 MB per 1080p frame) the largest PCIe / host-memory stream of the whole stage and caps the 8-GPU end-to-end rate.
// K9a packs, per frame, only the blocks that b2_coef_present() marks, macroblocks in raster order, blocks in index
// order (the host walks the stream with the same rule, include/b2enc_types.h).  K9b, on the copy-out stream, writes
// exactly the used bytes into pinned host memory with coalesced 16-byte stores (the size is only known on the
// device, so a cudaMemcpy cannot do it without a host round trip).
// Bound: latency (one CTA per frame does a block-wide prefix sum); algorithmic bytes = 4 B/MB of masks + 2 x the
// packed size.
#include "b2_common.cuh"
#include "b2_internal.h"

namespace {

constexpr int K9_THREADS = 1024;

__global__ void __launch_bounds__(K9_THREADS)
k9a_pack_levels_kernel(const b2_mbinfo_t *__restrict__ info, const b2_mbcoef_t *__restrict__ coef, uint8_t *__restrict__ packed,
                       size_t packed_stride, uint32_t *__restrict__ nblocks, unsigned long long *__restrict__ cum_bytes, int nmb)
{
    __shared__ uint32_t s_warp[32];
    const int frame = blockIdx.x;
    const b2_mbinfo_t *fi = info + (size_t)frame * nmb;
    const b2_mbcoef_t *fc = coef + (size_t)frame * nmb;
    uint4 *out = (uint4 *)(packed + (size_t)frame * packed_stride);
    const int per = (nmb + K9_THREADS - 1) / K9_THREADS;
    const int mb0 = threadIdx.x * per, mb1 = min(nmb, mb0 + per);
    uint32_t cnt = 0;
    for (int m = mb0; m < mb1; m++) cnt += __popc(b2_coef_present(&fi[m]));
    // block-wide exclusive scan of cnt
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
        s_warp[lane] = wi - w;                                        // exclusive offset of each warp
        if (lane == 31) {
            nblocks[frame] = wi;
            atomicAdd(cum_bytes, (unsigned long long)wi * 32ull);
        }
    }
    __syncthreads();
    uint32_t off = s_warp[warp] + inc - cnt;                           // in 32-byte blocks
    for (int m = mb0; m < mb1; m++) {
        uint32_t pm = b2_coef_present(&fi[m]);
        const uint4 *src = (const uint4 *)fc[m].blk;
        while (pm) {
            const int b = __ffs(pm) - 1;
            pm &= pm - 1;
            out[2 * off] = src[2 * b]; out[2 * off + 1] = src[2 * b + 1];
            off++;
        }
    }
}

// copy the used part of each frame's packed stream (and its size) into pinned host memory; grid (ctas, frames)
__global__ void __launch_bounds__(256)
k9b_copy_out_kernel(const uint8_t *__restrict__ packed, size_t packed_stride, const uint32_t *__restrict__ nblocks,
                    uint8_t *__restrict__ h_packed, uint32_t *__restrict__ h_nblocks)
{
    const int frame = blockIdx.y;
    const uint32_t n16 = nblocks[frame] * 2u;                          // 16-byte words
    const uint4 *src = (const uint4 *)(packed + (size_t)frame * packed_stride);
    uint4 *dst = (uint4 *)(h_packed + (size_t)frame * packed_stride);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = src[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) h_nblocks[frame] = nblocks[frame];
}

}  // namespace

int b2_launch_pack_levels(const b2_mbinfo_t *d_info, const b2_mbcoef_t *d_coef, uint8_t *d_packed, size_t packed_stride,
                          uint32_t *d_nblocks, unsigned long long *d_cum_bytes, int nmb, int nframes, cudaStream_t st)
{
    k9a_pack_levels_kernel<<<nframes, K9_THREADS, 0, st>>>(d_info, d_coef, d_packed, packed_stride, d_nblocks, d_cum_bytes, nmb);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

int b2_launch_pack_copy_out(const uint8_t *d_packed, size_t packed_stride, const uint32_t *d_nblocks, uint8_t *h_packed,
                            uint32_t *h_nblocks, int nframes, cudaStream_t st)
{
    k9b_copy_out_kernel<<<dim3(8, nframes), 256, 0, st>>>(d_packed, packed_stride, d_nblocks, h_packed, h_nblocks);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}
