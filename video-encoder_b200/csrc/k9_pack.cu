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
// Bound: latency / HBM; algorithmic bytes = 4 B/MB of masks + 2 x the packed size.
//
// K9a is two small launches so that a frame is packed by many CTAs instead of one (one CTA walking 8,160 macroblocks was a
// 0.1-0.3 ms link in the per-frame dependency chain): k9a_count sums the present blocks of every chunk of 256 macroblocks,
// k9a_scatter adds up the chunks in front of its own (at most a few dozen numbers), then one warp per macroblock copies that
// macroblock's blocks -- lane = 16-byte half block -- to their place in the stream.
#include "b2_common.cuh"
#include "b2_internal.h"

namespace {

constexpr int K9_CHUNK = 256;             // macroblocks per CTA = threads per CTA

__global__ void __launch_bounds__(K9_CHUNK)
k9a_count_kernel(const b2_mbinfo_t *__restrict__ info, uint32_t *__restrict__ chunk_cnt, b2_mbinfo_packed_t *__restrict__ pinfo, int nmb, int nchunk)
{
    __shared__ uint32_t s_warp[K9_CHUNK / 32];
    const int frame = blockIdx.y, m = blockIdx.x * K9_CHUNK + threadIdx.x;
    uint32_t cnt = 0;
    if (m < nmb) {
        const b2_mbinfo_t mi = info[(size_t)frame * nmb + m];
        cnt = __popc(b2_coef_present(&mi));
        // the 24-byte record that crosses PCIe instead of the 48-byte one (include/b2enc_types.h)
        const b2_mbinfo_packed_t p = b2_mbinfo_pack(&mi);
        uint2 *dst = (uint2 *)&pinfo[(size_t)frame * nmb + m];
        dst[0] = make_uint2(p.w[0], p.w[1]); dst[1] = make_uint2(p.w[2], p.w[3]); dst[2] = make_uint2(p.w[4], p.w[5]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < K9_CHUNK / 32; i++) t += s_warp[i];
        chunk_cnt[(size_t)frame * nchunk + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(K9_CHUNK)
k9a_scatter_kernel(const b2_mbinfo_t *__restrict__ info, const b2_mbcoef_t *__restrict__ coef, const uint32_t *__restrict__ chunk_cnt,
                   uint8_t *__restrict__ packed, size_t packed_stride, uint32_t *__restrict__ nblocks,
                   unsigned long long *__restrict__ cum_bytes, int nmb, int nchunk)
{
    __shared__ uint32_t s_off[K9_CHUNK + 1];      // exclusive block offset of every macroblock of the chunk inside the frame's stream
    __shared__ uint32_t s_mask[K9_CHUNK];
    __shared__ uint32_t s_warp[K9_CHUNK / 32];
    __shared__ uint32_t s_base;
    const int frame = blockIdx.y, chunk = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const b2_mbinfo_t *fi = info + (size_t)frame * nmb;
    const b2_mbcoef_t *fc = coef + (size_t)frame * nmb;
    const int m = chunk * K9_CHUNK + threadIdx.x;
    const uint32_t pm = m < nmb ? b2_coef_present(&fi[m]) : 0;
    const uint32_t cnt = __popc(pm);
    s_mask[threadIdx.x] = pm;
    // blocks in front of this chunk: the chunk totals of k9a_count (a frame has a few dozen chunks)
    if (warp == 0) {
        uint32_t b = 0;
        for (int i = lane; i < nchunk; i += 32) {
            if (i < chunk) b += chunk_cnt[(size_t)frame * nchunk + i];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if (lane == 0) s_base = b;
    }
    // exclusive scan of cnt over the chunk
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int i = 0; i < warp; i++) woff += s_warp[i];
    s_off[threadIdx.x] = s_base + woff + inc - cnt;
    if (chunk == nchunk - 1 && threadIdx.x == K9_CHUNK - 1) {          // last thread of the last chunk knows the frame total
        const uint32_t total = s_base + woff + inc;
        nblocks[frame] = total;
        atomicAdd(cum_bytes, (unsigned long long)total * 32ull);
    }
    __syncthreads();
    // copy: one warp per macroblock, lane = (block, 16-byte half); a macroblock rarely holds more than 16 present blocks
    uint4 *out = (uint4 *)(packed + (size_t)frame * packed_stride);
    for (int j = warp; j < K9_CHUNK; j += K9_CHUNK / 32) {
        const int mm = chunk * K9_CHUNK + j;
        if (mm >= nmb) break;
        const uint32_t mask = s_mask[j];
        const int nb = __popc(mask);
        if (!nb) continue;
        const uint4 *src = (const uint4 *)fc[mm].blk;
        const uint32_t off = s_off[j];
        for (int t0 = 0; t0 < nb; t0 += 16) {                           // 16 present blocks per pass
            const int t = t0 + (lane >> 1);
            if (t < nb) {
                const int bsel = (int)__fns(mask, 0, t + 1);            // the t-th present block
                out[2 * (off + t) + (lane & 1)] = src[2 * bsel + (lane & 1)];
            }
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
                          uint32_t *d_nblocks, unsigned long long *d_cum_bytes, uint32_t *d_chunk_cnt, b2_mbinfo_packed_t *d_pinfo, int nmb,
                          int nframes, cudaStream_t st)
{
    const int nchunk = b2_pack_chunks(nmb);
    k9a_count_kernel<<<dim3(nchunk, nframes), K9_CHUNK, 0, st>>>(d_info, d_chunk_cnt, d_pinfo, nmb, nchunk);
    k9a_scatter_kernel<<<dim3(nchunk, nframes), K9_CHUNK, 0, st>>>(d_info, d_coef, d_chunk_cnt, d_packed, packed_stride, d_nblocks, d_cum_bytes,
                                                                   nmb, nchunk);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

int b2_pack_chunks(int nmb) { return (nmb + K9_CHUNK - 1) / K9_CHUNK; }

int b2_launch_pack_copy_out(const uint8_t *d_packed, size_t packed_stride, const uint32_t *d_nblocks, uint8_t *h_packed,
                            uint32_t *h_nblocks, int nframes, cudaStream_t st)
{
    k9b_copy_out_kernel<<<dim3(8, nframes), 256, 0, st>>>(d_packed, packed_stride, d_nblocks, h_packed, h_nblocks);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}
