/*
 * b2h_picture.h -- registry of the pictures handed out by b2_picture_alloc (the mirror of x264_picture_alloc,
 * av_encode.c:415).  Internal to libb2enc.so; not part of the C-ABI.
 *
 * The reference converts every decoded frame INTO that one picture (sws_scale, av_encode.c:545-547) and then hands the
 * picture to the encoder (:970).  Both calls are ours, so the picture record is where they meet: when b2_sws_scale is
 * given the planes of a registered picture as destination it only stages the raw source in page-locked memory
 * ("deferred conversion") and b2_encoder_encode uploads that once -- K0 then converts it straight into the encoder's
 * device planes.  Without the record the same frame would cross PCIe three times (up, down, up again).
 */
#ifndef B2H_PICTURE_H
#define B2H_PICTURE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2h_picrec {
    uint8_t *base;              /* the page-locked allocation behind img.plane[0..2]                          */
    size_t bytes;
    int width, height;
    /* deferred conversion, written by b2_sws_scale and consumed by b2_encoder_encode.  Two staging buffers alternate so that
     * the upload of one picture (asynchronous DMA issued by b2_encoder_encode) overlaps the staging of the next: the caller's
     * thread never waits for PCIe.  busy_eng / busy_ticket: the upload that still reads a buffer (b2_engine_put_wait). */
    int deferred;               /* 1: stage[cur] holds the raw source of format `fmt`; the I420 planes are stale */
    int fmt;                    /* B2_FMT_*                                                                      */
    int cur;                    /* buffer the last b2_sws_scale filled                                           */
    uint8_t *stage[2];          /* page-locked, tight planes one after the other (b2_fmt_layout)                 */
    size_t stage_bytes[2];
    void *busy_eng[2];
    long busy_ticket[2];
    struct b2h_picrec *next;
} b2h_picrec_t;

/* record of the picture whose first plane starts at `plane0`, or NULL */
b2h_picrec_t *b2h_picture_find(const uint8_t *plane0);
/* the staging buffer to fill next: page-locked, at least `bytes`, no upload reading it any more; NULL on failure.  Makes it
 * r->cur. */
uint8_t *b2h_picture_stage(b2h_picrec_t *r, size_t bytes);
/* forget uploads issued through engine `eng` (it is about to be destroyed; its streams have been drained) */
void b2h_picture_forget_engine(void *eng);

void *b2_pinned_alloc(size_t n);
void b2_pinned_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
