// b2_device.cu -- device plumbing shared by all kernels: TMA descriptors over padded plane
// stacks and K6, the border-replication kernel (the device-side twin of what x264 does to
// its internal frames so that unrestricted motion vectors may point outside the picture).
#include "b2_common.cuh"
#include "b2_internal.h"

// ---- TMA descriptor ------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled()
{
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || p == nullptr) {
        fprintf(stderr, "b2enc: cuTensorMapEncodeTiled not available from the driver\n");
        return nullptr;
    }
    fn = (PFN_encodeTiled)p;
    return fn;
}

int b2_make_plane_tmap(CUtensorMap *tm, const void *base, int pitch, int rows, int nplanes, int bw, int bh)
{
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return -1;
    cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)nplanes};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "b2enc: cuTensorMapEncodeTiled failed (%d) pitch=%d rows=%d n=%d box=%dx%d\n", (int)r, pitch,
                rows, nplanes, bw, bh);
        return -1;
    }
    return 0;
}

// u32 plane stack with the geometry of a luma plane stack (the min | max block sums K1 prunes with); box {bw, 1, 1} words
int b2_make_plane_tmap32(CUtensorMap *tm, const void *base, int pitch, int rows, int nplanes, int bw)
{
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return -1;
    cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)nplanes};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * rows * 4};
    cuuint32_t box[3] = {(cuuint32_t)bw, 1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "b2enc: cuTensorMapEncodeTiled (u32) failed (%d) pitch=%d rows=%d n=%d box=%d\n", (int)r, pitch, rows, nplanes, bw);
        return -1;
    }
    return 0;
}

// ---- K6: border replication -------------------------------------------------------------------
// planes: [n][rows][pitch] with the picture origin at (pad,pad); every byte outside the
// iw x ih interior becomes interior[clamp(y)][clamp(x)].  One thread per aligned 32-bit word.
__global__ void k6_extend_border_kernel(uint8_t *planes, int pitch, int rows, int pad, int iw, int ih)
{
    const int wpr = pitch >> 2;
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (wx >= wpr) return;
    uint8_t *plane = planes + (size_t)blockIdx.z * pitch * rows;
    const int x0 = wx * 4 - pad, yy = y - pad;
    if (yy >= 0 && yy < ih && x0 >= 0 && x0 + 3 < iw) return;       // interior word
    const int cy = min(max(yy, 0), ih - 1);
    const uint8_t *src = plane + (size_t)(cy + pad) * pitch + pad;
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        int cx = min(max(x0 + b, 0), iw - 1);
        v |= (uint32_t)src[cx] << (8 * b);
    }
    *(uint32_t *)(plane + (size_t)y * pitch + wx * 4) = v;
}

int b2_launch_extend_border(uint8_t *d_planes, int pitch, int rows, int nplanes, int pad, int iw, int ih,
                            cudaStream_t st)
{
    dim3 block(128);
    dim3 grid((pitch / 4 + 127) / 128, rows, nplanes);
    k6_extend_border_kernel<<<grid, block, 0, st>>>(d_planes, pitch, rows, pad, iw, ih);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- K6 (frame form): borders of Y, U and V of every frame in ONE launch, visiting border bytes only -----------
// Per plane the border is flattened into a 1-D list of 8-byte units: first the 2*pad full-width rows above and below
// the picture, then the left+right pad units of the ih interior rows.  blockIdx.y = frame * 3 + plane.  Pads (64 / 32), the
// coded widths (multiples of 16 / 8) and the pitches are multiples of 8, so a unit never straddles the picture edge: it is
// either one replicated edge pixel (left / right of the picture) or an aligned 8-byte copy of the edge row (above / below).
__global__ void __launch_bounds__(128)
k6_extend_border_yuv_kernel(uint8_t *y, uint8_t *u, uint8_t *v, int pitch, int pitchc, size_t stride_y, size_t stride_c,
                            int w16, int h16)
{
    const int plane = blockIdx.y % 3, frame = blockIdx.y / 3;
    const int pt = plane ? pitchc : pitch, pad = plane ? B2_PADC : B2_PAD;
    const int iw = plane ? w16 >> 1 : w16, ih = plane ? h16 >> 1 : h16;
    uint8_t *base = (plane == 0 ? y + frame * stride_y : (plane == 1 ? u : v) + frame * stride_c);
    const int upr = pt >> 3, nfull = 2 * pad * upr;
    const int lu = pad >> 3, ru = (pt - pad - iw) >> 3, side = lu + ru;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int row, ux;
    if (t < nfull) {
        const int r = t / upr;
        ux = t - r * upr;
        row = r < pad ? r : ih + r;           // r in [pad,2pad) -> rows ih+pad .. ih+2pad-1
    } else {
        t -= nfull;
        const int r = t / side, k = t - r * side;
        if (r >= ih) return;
        row = pad + r;
        ux = k < lu ? k : ((pad + iw) >> 3) + (k - lu);
    }
    const int x0 = ux * 8 - pad, yy = row - pad;
    const int cy = min(max(yy, 0), ih - 1);
    const uint8_t *src = base + (size_t)(cy + pad) * pt + pad;
    uint2 w;
    if (x0 < 0 || x0 >= iw) {
        const uint32_t e = (uint32_t)src[x0 < 0 ? 0 : iw - 1] * 0x01010101u;
        w = make_uint2(e, e);
    } else {
        w = *(const uint2 *)(src + x0);
    }
    *(uint2 *)(base + (size_t)row * pt + ux * 8) = w;
}

int b2_launch_extend_border_yuv(uint8_t *d_y, uint8_t *d_u, uint8_t *d_v, int pitch, int rows, int pitchc, int rowsc,
                                size_t stride_y, size_t stride_c, int w16, int h16, int nframes, cudaStream_t st)
{
    (void)rows; (void)rowsc;
    if ((pitch | pitchc) & 7) { fprintf(stderr, "b2enc: plane pitches must be multiples of 8\n"); return -1; }
    const int units = 2 * B2_PAD * (pitch / 8) + h16 * ((pitch - w16) / 8);      // luma is the largest plane
    dim3 block(128);
    dim3 grid((units + 127) / 128, 3 * nframes);
    k6_extend_border_yuv_kernel<<<grid, block, 0, st>>>(d_y, d_u, d_v, pitch, pitchc, stride_y, stride_c, w16, h16);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}
