/*
 * b2h_sws.c -- host half of the drop-in for the reference's use of libswscale (av_encode.c:427-430, :441, :545-547):
 * same-size conversion of a decoder picture into I420, executed on the GPU by kernel K0.
 *
 * Two forms, chosen per call by where the destination planes live:
 *  * destination = a picture from b2_picture_alloc (what the reference always passes: x264.pic_in, av_encode.c:415,
 *    :545-547).  That picture only exists to be handed to b2_encoder_encode next (:970), so the conversion is DEFERRED:
 *    the source is copied once into page-locked staging that belongs to the picture (b2h_picture.h), b2_encoder_encode
 *    uploads it once, and K0 converts it straight into the encoder's device planes.  The frame crosses PCIe once instead
 *    of three times (up, down, up again).  The picture's own planes are NOT written in this form (nothing in the reference
 *    reads them; create the context with B2_SWS_HOST_OUTPUT to get them).  The staging is double buffered and the copy into
 *    it is split over a few helper threads for large pictures, so the caller's thread neither waits for PCIe nor spends a
 *    whole memcpy of a 12 MB 4K picture per frame.
 *  * any other destination: synchronous host -> GPU -> host round trip (csrc/b2_sws.cu), sws_scale's host-out contract.
 * In both forms the source has been read when b2_sws_scale returns (the reference frees it right away, av_encode.c:550).
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "b2enc.h"
#include "b2h_picture.h"

/* ---- staging copy, split by rows over the caller and a few helper threads --------------------------------------------- */
#define SWS_MAX_HELPERS 3
typedef struct {
    uint8_t *dst; const uint8_t *src; size_t dst_pitch, src_pitch, row_bytes; int rows;
} copy_job_t;
typedef struct {
    pthread_t th[SWS_MAX_HELPERS];
    int n;                                  /* helper threads running */
    pthread_mutex_t mu;
    pthread_cond_t cv_go, cv_done;
    copy_job_t job[SWS_MAX_HELPERS];
    unsigned long gen;                      /* bumped when a new set of jobs is posted */
    int pending, stop;
} copy_pool_t;

static void copy_rows(const copy_job_t *j)
{
    if (j->dst_pitch == j->row_bytes && j->src_pitch == j->row_bytes) { memcpy(j->dst, j->src, j->row_bytes * (size_t)j->rows); return; }
    for (int y = 0; y < j->rows; y++) memcpy(j->dst + (size_t)y * j->dst_pitch, j->src + (size_t)y * j->src_pitch, j->row_bytes);
}

typedef struct { copy_pool_t *pool; int idx; } helper_arg_t;
static void *copy_helper(void *arg)
{
    helper_arg_t *ha = (helper_arg_t *)arg;
    copy_pool_t *p = ha->pool;
    const int me = ha->idx;
    free(ha);
    unsigned long seen = 0;
    pthread_mutex_lock(&p->mu);
    for (;;) {
        while (!p->stop && p->gen == seen) pthread_cond_wait(&p->cv_go, &p->mu);
        if (p->stop) break;
        seen = p->gen;
        const copy_job_t j = p->job[me];
        pthread_mutex_unlock(&p->mu);
        if (j.rows > 0) copy_rows(&j);
        pthread_mutex_lock(&p->mu);
        if (--p->pending == 0) pthread_cond_signal(&p->cv_done);
    }
    pthread_mutex_unlock(&p->mu);
    return NULL;
}

static copy_pool_t *copy_pool_create(int helpers)
{
    copy_pool_t *p = (copy_pool_t *)calloc(1, sizeof(*p));
    if (!p) return NULL;
    pthread_mutex_init(&p->mu, NULL); pthread_cond_init(&p->cv_go, NULL); pthread_cond_init(&p->cv_done, NULL);
    for (int i = 0; i < helpers && i < SWS_MAX_HELPERS; i++) {
        helper_arg_t *ha = (helper_arg_t *)malloc(sizeof(*ha));
        if (!ha) break;
        ha->pool = p; ha->idx = i;
        if (pthread_create(&p->th[i], NULL, copy_helper, ha)) { free(ha); break; }
        p->n++;
    }
    return p;
}

static void copy_pool_destroy(copy_pool_t *p)
{
    if (!p) return;
    pthread_mutex_lock(&p->mu);
    p->stop = 1;
    pthread_cond_broadcast(&p->cv_go);
    pthread_mutex_unlock(&p->mu);
    for (int i = 0; i < p->n; i++) pthread_join(p->th[i], NULL);
    pthread_mutex_destroy(&p->mu); pthread_cond_destroy(&p->cv_go); pthread_cond_destroy(&p->cv_done);
    free(p);
}

/* one plane: rows split evenly over the helpers and the calling thread */
static void copy_plane(copy_pool_t *p, uint8_t *dst, size_t dst_pitch, const uint8_t *src, size_t src_pitch, size_t row_bytes, int rows)
{
    const int parts = p && p->n > 0 && row_bytes * (size_t)rows >= ((size_t)1 << 20) ? p->n + 1 : 1;
    copy_job_t mine = {dst, src, dst_pitch, src_pitch, row_bytes, rows};
    if (parts == 1) { copy_rows(&mine); return; }
    const int per = (rows + parts - 1) / parts;
    pthread_mutex_lock(&p->mu);
    for (int i = 0; i < p->n; i++) {
        const int r0 = (i + 1) * per, r1 = r0 + per < rows ? r0 + per : rows;
        copy_job_t j = {dst + (size_t)r0 * dst_pitch, src + (size_t)r0 * src_pitch, dst_pitch, src_pitch, row_bytes, r1 > r0 ? r1 - r0 : 0};
        p->job[i] = j;
    }
    p->pending = p->n;
    p->gen++;
    pthread_cond_broadcast(&p->cv_go);
    pthread_mutex_unlock(&p->mu);
    mine.rows = per < rows ? per : rows;
    copy_rows(&mine);
    pthread_mutex_lock(&p->mu);
    while (p->pending > 0) pthread_cond_wait(&p->cv_done, &p->mu);
    pthread_mutex_unlock(&p->mu);
}

void *b2_sws_rt_create(int w, int h, int fmt);
void b2_sws_rt_free(void *rt);
int b2_sws_rt_scale(void *rt, const uint8_t *const src[], const int srcStride[], uint8_t *const dst[], const int dstStride[]);

struct b2_sws_context {
    int w, h, fmt, host_output;
    size_t in_bytes;
    void *rt;                   /* GPU round-trip state (device, buffers, stream) */
    copy_pool_t *pool;          /* helper threads of the staging copy (created with the first large picture) */
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
    copy_pool_destroy(c->pool);
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
        /* deferred form: stage the raw source with the picture (double buffered: b2h_picture_stage waits for the upload of the
         * picture before last, not for the last one) */
        uint8_t *p = b2h_picture_stage(rec, c->in_bytes);
        if (!p) { fprintf(stderr, "b2enc: b2_sws_scale: cannot allocate page-locked staging\n"); return -1; }
        if (!c->pool && c->in_bytes >= ((size_t)2 << 20)) {
            long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
            const char *e = getenv("B2ENC_SWS_THREADS");
            int helpers = e ? atoi(e) - 1 : (c->in_bytes >= ((size_t)8 << 20) ? 3 : 2);
            if (ncpu > 0 && helpers > ncpu / 4) helpers = (int)(ncpu / 4);
            c->pool = copy_pool_create(helpers < 0 ? 0 : helpers);
        }
        int rb[3], rws[3];
        const int np = b2_fmt_layout(c->fmt, w, h, rb, rws);
        for (int k = 0; k < np; k++) {
            copy_plane(c->pool, p, (size_t)rb[k], src[k], (size_t)srcStride[k], (size_t)rb[k], rws[k]);
            p += (size_t)rb[k] * rws[k];
        }
        rec->fmt = c->fmt;
        rec->deferred = 1;
        return h;
    }
    if (rec) rec->deferred = 0;                           /* its planes are about to hold a converted frame */
    return b2_sws_rt_scale(c->rt, src, srcStride, dst, dstStride);
}
