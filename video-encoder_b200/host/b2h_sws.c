/*
 * b2h_sws.c -- host half of the drop-in for the reference's use of libswscale (av_encode.c:427-430, :441, :545-547):
 * same-size conversion of a decoder picture into I420, executed on the GPU by kernel K0.
 *
 * Two forms, chosen per call by where the destination planes live:
 *  * destination = a picture from b2_picture_alloc (what the reference always passes: x264.pic_in, av_encode.c:415,
 *    :545-547).  That picture only exists to be handed to b2_encoder_encode next (:970), so the conversion is DEFERRED:
 *    the source is copied once into page-locked staging that belongs to the picture (b2h_picture.h), b2_encoder_encode
 *    uploads it once, and K0 converts it straight into the encoder's device planes.  The frame crosses PCIe once instead
 *    of three times (up, down, up again).  For yuv420p sources the staging IS the picture (the conversion is a plane
 *    copy), so its planes are valid on return; for the other formats they are not written unless the context was created
 *    with B2_SWS_HOST_OUTPUT.
 *  * any other destination: synchronous host -> GPU -> host round trip (csrc/b2_sws.cu), sws_scale's host-out contract.
 * In both forms the source has been read when b2_sws_scale returns (the reference frees it right away, av_encode.c:550).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b2enc.h"
#include "b2h_picture.h"

void *b2_sws_rt_create(int w, int h, int fmt);
void b2_sws_rt_free(void *rt);
int b2_sws_rt_scale(void *rt, const uint8_t *const src[], const int srcStride[], uint8_t *const dst[], const int dstStride[]);

struct b2_sws_context {
    int w, h, fmt, host_output;
    size_t in_bytes;
    void *rt;                   /* GPU round-trip state (device, buffers, stream) */
};

b2_sws_context_t *b2_sws_getContext(int srcW, int srcH, int srcFormat, int dstW, int dstH, int dstFormat, int flags,
                                    void *srcFilter, void *dstFilter, const double *param)
{
    (void)srcFilter; (void)dstFilter; (void)param;
    if (srcW != dstW || srcH != dstH || dstFormat != B2_FMT_YUV420P || srcW < 2 || srcH < 2) {
        fprintf(stderr, "b2enc: b2_sws_getContext supports same-size conversion to yuv420p only\n");
        return NULL;
    }
    int rb[3], rws[3];
    if (!b2_fmt_layout(srcFormat, srcW, srcH, rb, rws)) {
        fprintf(stderr, "b2enc: unsupported source pixel format %d\n", srcFormat);
        return NULL;
    }
    if (!b2_fmt_size_ok(srcFormat, srcW, srcH)) {
        fprintf(stderr, "b2enc: source format %d cannot be converted at %dx%d (see b2_fmt_size_ok in b2enc_types.h)\n", srcFormat, srcW, srcH);
        return NULL;
    }
    b2_sws_context_t *c = (b2_sws_context_t *)calloc(1, sizeof(*c));
    if (!c) return NULL;
    c->w = srcW; c->h = srcH; c->fmt = srcFormat; c->host_output = (flags & B2_SWS_HOST_OUTPUT) != 0;
    for (int p = 0; p < 3; p++) c->in_bytes += (size_t)rb[p] * rws[p];
    c->rt = b2_sws_rt_create(srcW, srcH, srcFormat);     /* NULL without a CUDA device: there is no CPU conversion */
    if (!c->rt) { free(c); return NULL; }
    return c;
}

void b2_sws_freeContext(b2_sws_context_t *c)
{
    if (!c) return;
    b2_sws_rt_free(c->rt);
    free(c);
}

int b2_sws_scale(b2_sws_context_t *c, const uint8_t *const src[], const int srcStride[], int srcSliceY, int srcSliceH,
                 uint8_t *const dst[], const int dstStride[])
{
    if (!c || !src || !dst || srcSliceY != 0 || srcSliceH != c->h) {
        fprintf(stderr, "b2enc: b2_sws_scale converts whole frames only (srcSliceY=0, srcSliceH=height)\n");
        return -1;
    }
    const int w = c->w, h = c->h, cw = (w + 1) / 2;
    b2h_picrec_t *rec = b2h_picture_find(dst[0]);
    if (rec && !c->host_output && rec->width == w && rec->height == h && dstStride[0] == w && dstStride[1] == cw && dstStride[2] == cw) {
        /* deferred form: stage the raw source with the picture; yuv420p: the picture's own planes are the staging */
        uint8_t *p = c->fmt == B2_FMT_YUV420P ? rec->base : b2h_picture_stage(rec, c->in_bytes);
        if (!p) { fprintf(stderr, "b2enc: b2_sws_scale: cannot allocate page-locked staging\n"); return -1; }
        int rb[3], rws[3];
        const int np = b2_fmt_layout(c->fmt, w, h, rb, rws);
        for (int k = 0; k < np; k++) {
            if (srcStride[k] == rb[k]) memcpy(p, src[k], (size_t)rb[k] * rws[k]);
            else for (int y = 0; y < rws[k]; y++) memcpy(p + (size_t)y * rb[k], src[k] + (size_t)y * srcStride[k], rb[k]);
            p += (size_t)rb[k] * rws[k];
        }
        rec->fmt = c->fmt;
        rec->deferred = c->fmt != B2_FMT_YUV420P;
        return h;
    }
    if (rec) rec->deferred = 0;                           /* its planes are about to hold a converted frame */
    return b2_sws_rt_scale(c->rt, src, srcStride, dst, dstStride);
}
