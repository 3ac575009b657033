// b2_sws.cu -- drop-in for the reference's use of libswscale (av_encode.c:427-430, :441, :545-547):
// same-size conversion of a decoder picture into I420, executed on the GPU by kernel K0.
// Synchronous like sws_scale: when b2_sws_scale returns the source has been read (the reference frees
// it right away, av_encode.c:550) and the destination planes are written.
#include <string.h>
#include "b2_common.cuh"
#include "b2_internal.h"
#include "../../include/b2enc.h"

struct b2_sws_context {
    int w, h, fmt, w16, h16, pitch, rows, pitchc, rowsc;
    size_t in_bytes;
    uint8_t *h_in = nullptr, *d_in = nullptr, *d_y = nullptr, *d_u = nullptr, *d_v = nullptr;
    cudaStream_t st = nullptr;
};

extern "C" void *b2_pinned_alloc(size_t n)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, n, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void b2_pinned_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" b2_sws_context_t *b2_sws_getContext(int srcW, int srcH, int srcFormat, int dstW, int dstH, int dstFormat, int flags,
                                               void *srcFilter, void *dstFilter, const double *param)
{
    (void)flags; (void)srcFilter; (void)dstFilter; (void)param;
    if (srcW != dstW || srcH != dstH || dstFormat != B2_FMT_YUV420P || srcW < 2 || srcH < 2) {
        fprintf(stderr, "b2enc: b2_sws_getContext supports same-size conversion to yuv420p only\n");
        return nullptr;
    }
    int rb[3], rws[3];
    if (!b2_fmt_layout(srcFormat, srcW, srcH, rb, rws)) {
        fprintf(stderr, "b2enc: unsupported source pixel format %d\n", srcFormat);
        return nullptr;
    }
    if (!b2_fmt_size_ok(srcFormat, srcW, srcH)) {
        fprintf(stderr, "b2enc: source format %d cannot be converted at %dx%d (see b2_fmt_size_ok in b2enc_types.h)\n", srcFormat, srcW, srcH);
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        fprintf(stderr, "b2enc: no CUDA device; conversion has no CPU fallback\n");
        return nullptr;
    }
    b2_sws_context *c = new b2_sws_context();
    c->w = srcW; c->h = srcH; c->fmt = srcFormat;
    c->w16 = (srcW + 15) & ~15; c->h16 = (srcH + 15) & ~15;
    c->pitch = c->w16 + 2 * B2_PAD; c->rows = c->h16 + 2 * B2_PAD;
    c->pitchc = (c->w16 / 2 + 2 * B2_PADC + 15) & ~15; c->rowsc = c->h16 / 2 + 2 * B2_PADC;
    c->in_bytes = 0;
    for (int p = 0; p < 3; p++) c->in_bytes += (size_t)rb[p] * rws[p];
    bool ok = cudaHostAlloc(&c->h_in, c->in_bytes, cudaHostAllocDefault) == cudaSuccess &&
              cudaMalloc(&c->d_in, c->in_bytes) == cudaSuccess && cudaMalloc(&c->d_y, (size_t)c->pitch * c->rows) == cudaSuccess &&
              cudaMalloc(&c->d_u, (size_t)c->pitchc * c->rowsc) == cudaSuccess &&
              cudaMalloc(&c->d_v, (size_t)c->pitchc * c->rowsc) == cudaSuccess && cudaStreamCreate(&c->st) == cudaSuccess;
    if (!ok) { fprintf(stderr, "b2enc: b2_sws_getContext: allocation failed\n"); b2_sws_freeContext(c); return nullptr; }
    return c;
}

extern "C" void b2_sws_freeContext(b2_sws_context_t *c)
{
    if (!c) return;
    cudaFreeHost(c->h_in); cudaFree(c->d_in); cudaFree(c->d_y); cudaFree(c->d_u); cudaFree(c->d_v);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

extern "C" int b2_sws_scale(b2_sws_context_t *c, const uint8_t *const src[], const int srcStride[], int srcSliceY, int srcSliceH,
                            uint8_t *const dst[], const int dstStride[])
{
    if (!c || srcSliceY != 0 || srcSliceH != c->h) {
        fprintf(stderr, "b2enc: b2_sws_scale converts whole frames only (srcSliceY=0, srcSliceH=height)\n");
        return -1;
    }
    const int w = c->w, h = c->h, cw = (w + 1) / 2, ch = (h + 1) / 2;
    uint8_t *p = c->h_in;
    int rb[3], rws[3];
    const int np = b2_fmt_layout(c->fmt, w, h, rb, rws);
    for (int k = 0; k < np; k++) {                      // strided source planes -> tight pinned staging
        for (int y = 0; y < rws[k]; y++) memcpy(p + (size_t)y * rb[k], src[k] + (size_t)y * srcStride[k], rb[k]);
        p += (size_t)rb[k] * rws[k];
    }
    B2_CUDA_OK(cudaMemcpyAsync(c->d_in, c->h_in, c->in_bytes, cudaMemcpyHostToDevice, c->st));
    if (b2_launch_convert(c->fmt, c->d_in, c->in_bytes, c->d_y, c->d_u, c->d_v, c->pitch, c->pitchc, (size_t)c->pitch * c->rows,
                          (size_t)c->pitchc * c->rowsc, w, h, 1, c->st))
        return -1;
    B2_CUDA_OK(cudaMemcpy2DAsync(dst[0], dstStride[0], c->d_y + (size_t)B2_PAD * c->pitch + B2_PAD, c->pitch, w, h,
                                 cudaMemcpyDeviceToHost, c->st));
    B2_CUDA_OK(cudaMemcpy2DAsync(dst[1], dstStride[1], c->d_u + (size_t)B2_PADC * c->pitchc + B2_PADC, c->pitchc, cw, ch,
                                 cudaMemcpyDeviceToHost, c->st));
    B2_CUDA_OK(cudaMemcpy2DAsync(dst[2], dstStride[2], c->d_v + (size_t)B2_PADC * c->pitchc + B2_PADC, c->pitchc, cw, ch,
                                 cudaMemcpyDeviceToHost, c->st));
    B2_CUDA_OK(cudaStreamSynchronize(c->st));
    return h;
}
