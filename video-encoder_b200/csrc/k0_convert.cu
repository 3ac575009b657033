// k0_convert.cu -- K0: frame conversion to the encoder's padded I420 planes.
//
// Replaces sws_scale(src fmt -> YUV420P, same size, SWS_FAST_BILINEAR) at av_encode.c:545-547 plus
// the copy x264_encoder_encode makes of pic_in into its internal frame (av_encode.c:970): the raw
// picture (tight layout, resident in HBM after one H2D copy) is converted straight into the
// macroblock-aligned, border-padded planes the search kernels read.  Bit-exact against
// oracle/b2o_convert.c (pinned by live libswscale golden vectors) followed by b2o_frame_load.
//
// Row N4 adds bgr24 (bit-exact), rgb24, yuv422p and yuv411p (tolerance-pinned, see oracle/b2o_convert.c).
// Bound: HBM.  Algorithmic bytes per frame = input bytes (1.5*W*H for 4:2:0, 2*W*H for 4:2:2, 3*W*H for RGB)
// + 1.5*W16*H16 written.  Each thread produces 16 output pixels with 128-bit loads/stores when the
// row is 16-byte aligned (all BASELINE.json resolutions), byte-wise otherwise.
#include "b2_common.cuh"
#include "b2_internal.h"

namespace {

struct K0Args {
    const uint8_t *in;          // [nframes][in_bytes] tight raw pictures
    size_t in_stride;           // bytes between pictures
    uint8_t *y, *u, *v;         // padded plane stacks (pointing at allocation start)
    int pitch, pitchc;
    size_t stride_y, stride_c;
    int w, h, w16, h16;
    int fmt;
};

__device__ __forceinline__ uint8_t avg_r(int a, int b, int r) { return (uint8_t)((a + b + r) >> 1); }

// 16 output pixels of row `row` (0..h16-1 luma, then U rows, then V rows) starting at x0; returns the store address or nullptr
__device__ __forceinline__ uint8_t *k0_row16(const K0Args &a, int frame, int row, int x0, uint8_t px[16])
{
    const int ch16 = a.h16 >> 1, cw16 = a.w16 >> 1;
    const int cw = (a.w + 1) >> 1, chh = (a.h + 1) >> 1;
    const uint8_t *in = a.in + frame * a.in_stride;
    int plane, y;
    if (row < a.h16) { plane = 0; y = row; }
    else if (row < a.h16 + ch16) { plane = 1; y = row - a.h16; }
    else { plane = 2; y = row - a.h16 - ch16; }
    const int ow = plane ? cw16 : a.w16;              // coded width of this plane
    if (x0 >= ow) return nullptr;
    const int pw = plane ? cw : a.w, ph = plane ? chh : a.h;   // picture width/height of this plane
    const int sy = min(y, ph - 1);
    uint8_t *dst = plane == 0 ? a.y + frame * a.stride_y + (size_t)(B2_PAD + y) * a.pitch + B2_PAD + x0
                              : (plane == 1 ? a.u : a.v) + frame * a.stride_c + (size_t)(B2_PADC + y) * a.pitchc + B2_PADC + x0;

    if (a.fmt == B2_FMT_YUV420P || (a.fmt == B2_FMT_NV12 && plane == 0)) {
        const uint8_t *src = in + (plane == 0 ? 0 : (size_t)a.w * a.h + (plane == 2 ? (size_t)cw * chh : 0)) + (size_t)sy * pw;
        if (x0 + 16 <= pw && (((uintptr_t)(src + x0)) & 15) == 0) {
            *(uint4 *)px = *(const uint4 *)(src + x0);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) px[i] = src[min(x0 + i, pw - 1)];
        }
    } else if (a.fmt == B2_FMT_NV12) {
        const uint8_t *src = in + (size_t)a.w * a.h + (size_t)sy * (2 * cw) + (plane - 1);
#pragma unroll
        for (int i = 0; i < 16; i++) px[i] = src[2 * min(x0 + i, pw - 1)];
    } else if (a.fmt == B2_FMT_BGR24 || a.fmt == B2_FMT_RGB24) {
        // row N4: 15-bit BT.601 limited-range coefficients.  bgr24 = libswscale's dedicated converter (bit-exact:
        // truncating, chroma from the truncated 2x2 mean); rgb24 rounded, chroma from the 2x2 sum (tolerance-pinned)
        const bool exact = a.fmt == B2_FMT_BGR24;
        const int ro = exact ? 2 : 0, bo = 2 - ro;
        if (plane == 0) {
            const uint8_t *src = in + (size_t)sy * (3 * a.w);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint8_t *p = src + 3 * min(x0 + i, pw - 1);
                const int v = 8414 * p[ro] + 16519 * p[1] + 3208 * p[bo];
                px[i] = (uint8_t)(exact ? (v >> 15) + 16 : (v + (16 << 15) + (1 << 14)) >> 15);
            }
        } else {
            const uint8_t *l0 = in + (size_t)(2 * sy) * (3 * a.w), *l1 = in + (size_t)min(2 * sy + 1, a.h - 1) * (3 * a.w);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int x = min(x0 + i, pw - 1);
                const uint8_t *p0 = l0 + 6 * x, *p1 = l1 + 6 * x;
                int r = p0[ro] + p0[3 + ro] + p1[ro] + p1[3 + ro], g = p0[1] + p0[4] + p1[1] + p1[4];
                int b = p0[bo] + p0[3 + bo] + p1[bo] + p1[3 + bo];
                const int cr = plane == 1 ? -4865 : 14392, cg = plane == 1 ? -9528 : -12061, cb = plane == 1 ? 14392 : -2332;
                if (exact) { r >>= 2; g >>= 2; b >>= 2; px[i] = (uint8_t)(((cr * r + cg * g + cb * b) >> 15) + 128); }
                else px[i] = (uint8_t)((cr * r + cg * g + cb * b + (128 << 17) + (1 << 16)) >> 17);
            }
        }
    } else if (a.fmt == B2_FMT_YUV422P || a.fmt == B2_FMT_YUV411P) {
        // row N4: planar 4:2:2 / 4:1:1 (DV): luma copied, chroma = rounded mean of the two source lines; 4:1:1 is then
        // doubled horizontally (even samples copied, odd samples the rounded mean of their neighbours)
        const int scw = a.fmt == B2_FMT_YUV422P ? cw : (a.w + 3) >> 2;
        if (plane == 0) {
            const uint8_t *src = in + (size_t)sy * a.w;
            if (x0 + 16 <= pw && (((uintptr_t)(src + x0)) & 15) == 0) {
                *(uint4 *)px = *(const uint4 *)(src + x0);
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) px[i] = src[min(x0 + i, pw - 1)];
            }
        } else {
            const uint8_t *pl = in + (size_t)a.w * a.h + (plane == 2 ? (size_t)scw * a.h : 0);
            const uint8_t *l0 = pl + (size_t)(2 * sy) * scw, *l1 = pl + (size_t)min(2 * sy + 1, a.h - 1) * scw;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int x = min(x0 + i, pw - 1);
                if (a.fmt == B2_FMT_YUV422P) px[i] = avg_r(l0[x], l1[x], 1);
                else {
                    const int k = x >> 1, k1 = min(k + 1, scw - 1);
                    const int va = (l0[k] + l1[k] + 1) >> 1, vb = (l0[k1] + l1[k1] + 1) >> 1;
                    px[i] = (uint8_t)((x & 1) ? (va + vb + 1) >> 1 : va);
                }
            }
        }
    } else {                                           // packed 4:2:2
        const int yo = a.fmt == B2_FMT_YUYV422 ? 0 : 1, uo = a.fmt == B2_FMT_YUYV422 ? 1 : 0;
        if (plane == 0) {
            const uint8_t *src = in + (size_t)sy * (2 * a.w) + yo;
#pragma unroll
            for (int i = 0; i < 16; i++) px[i] = src[2 * min(x0 + i, pw - 1)];
        } else {
            const int co = uo + 2 * (plane - 1);
            const uint8_t *l0 = in + (size_t)(2 * sy) * (2 * a.w) + co, *l1 = l0 + 2 * a.w;
            const int body = cw & ~7;                 // libswscale 9.1.100: SIMD body rounds, scalar tail truncates
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int x = min(x0 + i, pw - 1);
                px[i] = avg_r(l0[4 * x], l1[4 * x], x < body ? 1 : 0);
            }
        }
    }
    return dst;
}

// K0_ROWS rows per thread: all loads of the four rows are in flight before the first store (memory-level parallelism; one row per
// thread left the kernel at 0.61 of the measured HBM rate).  Row groups never straddle planes: h16 is a multiple of 16.
constexpr int K0_ROWS = 4;

__global__ void __launch_bounds__(128)
k0_convert_kernel(K0Args a)
{
    const int frame = blockIdx.z;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    __align__(16) uint8_t px[K0_ROWS][16];
    uint8_t *dst[K0_ROWS];
#pragma unroll
    for (int r = 0; r < K0_ROWS; r++) dst[r] = k0_row16(a, frame, blockIdx.y * K0_ROWS + r, x0, px[r]);
#pragma unroll
    for (int r = 0; r < K0_ROWS; r++)
        if (dst[r]) *(uint4 *)dst[r] = *(const uint4 *)px[r];
}

}  // namespace

int b2_launch_convert(int fmt, const uint8_t *d_in, size_t in_stride, uint8_t *d_y, uint8_t *d_u, uint8_t *d_v, int pitch,
                      int pitchc, size_t stride_y, size_t stride_c, int w, int h, int nframes, cudaStream_t st)
{
    K0Args a;
    a.in = d_in; a.in_stride = in_stride; a.y = d_y; a.u = d_u; a.v = d_v; a.pitch = pitch; a.pitchc = pitchc;
    a.stride_y = stride_y; a.stride_c = stride_c; a.w = w; a.h = h; a.w16 = (w + 15) & ~15; a.h16 = (h + 15) & ~15; a.fmt = fmt;
    if (fmt < 0 || fmt >= B2_FMT_COUNT || !b2_fmt_size_ok(fmt, w, h)) {
        fprintf(stderr, "b2enc: unsupported raw format / size for conversion (format %d, %dx%d; see b2_fmt_size_ok)\n", fmt, w, h);
        return -1;
    }
    dim3 block(128);
    dim3 grid((a.w16 / 16 + 127) / 128, (a.h16 + 2 * (a.h16 / 2)) / K0_ROWS, nframes);
    k0_convert_kernel<<<grid, block, 0, st>>>(a);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}
