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
#include <time.h>
#include <unistd.h>
#include "b2enc.h"
#include "b2h_picture.h"

/* ---- staging copy, shared between the caller and a few helper threads -------------------------------------------------- */
/* The copy is what the caller's thread spends most of its time on (B2ENC_STATS: it waits ~1 % of the time for the encoder), and the
 * helpers share the host's cores with the entropy workers, so a fixed split would make the caller wait for whichever helper got
 * its time slice last.  Instead a picture is cut into chunks of rows (~64 KB) that the caller and the helpers CLAIM one at a time:
 * a helper that comes late simply finds fewer chunks left, and the caller only ever waits for chunks that are already being copied.
 * One post per picture; a helper that has just finished keeps looking for the next picture for ~0.1 ms before it sleeps on the
 * condition variable -- a producer faster than ~5,000 pictures/s never pays a futex wake-up, a real-time one costs the helpers
 * 0.1 ms of polling per picture. */
#define SWS_MAX_HELPERS 3
#define SWS_MAX_CHUNKS 256
#define SWS_CHUNK_BYTES ((size_t)64 << 10)
#define SWS_SPIN 2000                       /* pause instructions (~40-140 cycles each) before sleeping; B2ENC_SWS_SPIN overrides */
static int sws_spin = SWS_SPIN;
typedef struct {
    uint8_t *dst; const uint8_t *src; size_t dst_pitch, src_pitch, row_bytes; int rows;
} copy_job_t;
typedef struct {
    pthread_t th[SWS_MAX_HELPERS];
    int n;                                  /* helper threads running */
    pthread_mutex_t mu;
    pthread_cond_t cv_go, cv_done;
    copy_job_t chunk[SWS_MAX_CHUNKS];       /* written by the poster before `ticket` announces the picture */
    uint64_t ticket;                        /* picture number << 32 | chunks of the picture << 16 | next unclaimed chunk.  Atomic; a chunk is
                                             * claimed by compare-and-swap on the whole word, so a thread that is still looking at an older
                                             * picture can never take a chunk of -- or read the chunk count of -- the current one */
    int done;                               /* chunks finished (atomic) */
    int stop;
} copy_pool_t;

static inline void cpu_relax(void)
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    __asm__ volatile("" ::: "memory");
#endif
}

static void copy_rows(const copy_job_t *j)
{
    if (j->rows <= 0) return;
    if (j->dst_pitch == j->row_bytes && j->src_pitch == j->row_bytes) { memcpy(j->dst, j->src, j->row_bytes * (size_t)j->rows); return; }
    for (int y = 0; y < j->rows; y++) memcpy(j->dst + (size_t)y * j->dst_pitch, j->src + (size_t)y * j->src_pitch, j->row_bytes);
}

/* claim and copy chunks of picture `pic` until none is left (or the pool has moved on); returns the number copied */
static int copy_claim(copy_pool_t *p, uint64_t pic)
{
    int n = 0;
    for (;;) {
        uint64_t t = __atomic_load_n(&p->ticket, __ATOMIC_ACQUIRE);
        if ((t >> 32) != pic) break;
        const int i = (int)(t & 0xffff), total = (int)((t >> 16) & 0xffff);
        if (i >= total) break;
        if (!__atomic_compare_exchange_n(&p->ticket, &t, t + 1, 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) continue;
        copy_rows(&p->chunk[i]);                          /* stable: the poster rewrites the table only after all `total` chunks are done */
        n++;
        if (__atomic_add_fetch(&p->done, 1, __ATOMIC_ACQ_REL) == total) {
            pthread_mutex_lock(&p->mu);                   /* the poster checks `done` under the lock before it sleeps */
            pthread_cond_signal(&p->cv_done);
            pthread_mutex_unlock(&p->mu);
        }
    }
    return n;
}

static void *copy_helper(void *arg)
{
    copy_pool_t *p = (copy_pool_t *)arg;
    uint64_t seen = 0;                                    /* number of the last picture this helper looked at */
    for (;;) {
        for (int i = 0; i < sws_spin && (__atomic_load_n(&p->ticket, __ATOMIC_ACQUIRE) >> 32) == seen; i++) cpu_relax();
        if ((__atomic_load_n(&p->ticket, __ATOMIC_ACQUIRE) >> 32) == seen) {
            pthread_mutex_lock(&p->mu);
            while (!p->stop && (__atomic_load_n(&p->ticket, __ATOMIC_ACQUIRE) >> 32) == seen) pthread_cond_wait(&p->cv_go, &p->mu);
            const int stop = p->stop;
            pthread_mutex_unlock(&p->mu);
            if (stop) break;
        }
        seen = __atomic_load_n(&p->ticket, __ATOMIC_ACQUIRE) >> 32;
        copy_claim(p, seen);
    }
    return NULL;
}

static copy_pool_t *copy_pool_create(int helpers)
{
    copy_pool_t *p = (copy_pool_t *)calloc(1, sizeof(*p));
    if (!p) return NULL;
    pthread_mutex_init(&p->mu, NULL); pthread_cond_init(&p->cv_go, NULL); pthread_cond_init(&p->cv_done, NULL);
    for (int i = 0; i < helpers && i < SWS_MAX_HELPERS; i++) {
        if (pthread_create(&p->th[i], NULL, copy_helper, p)) break;
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

/* a picture of np planes, copied by the calling thread and whichever helpers get to it */
static void copy_planes(copy_pool_t *p, int np, uint8_t *const dst[], const size_t dst_pitch[], const uint8_t *const src[], const size_t src_pitch[],
                        const size_t row_bytes[], const int rows[])
{
    size_t total = 0;
    for (int k = 0; k < np; k++) total += row_bytes[k] * (size_t)rows[k];
    if (!p || p->n == 0 || total < ((size_t)1 << 20)) {
        for (int k = 0; k < np; k++) { const copy_job_t j = {dst[k], src[k], dst_pitch[k], src_pitch[k], row_bytes[k], rows[k]}; copy_rows(&j); }
        return;
    }
    /* chunks of whole rows, ~64 KB each but never more than the table holds */
    size_t chunk_bytes = SWS_CHUNK_BYTES;
    if (total / chunk_bytes > SWS_MAX_CHUNKS - 8) chunk_bytes = total / (SWS_MAX_CHUNKS - 8);
    int n = 0;
    for (int k = 0; k < np; k++) {
        int per = (int)(chunk_bytes / (row_bytes[k] ? row_bytes[k] : 1));
        if (per < 1) per = 1;
        while ((rows[k] + per - 1) / per > SWS_MAX_CHUNKS - n - (np - 1 - k)) per++;        /* very wide rows: still fits the table */
        for (int r0 = 0; r0 < rows[k]; r0 += per) {
            const copy_job_t j = {dst[k] + (size_t)r0 * dst_pitch[k], src[k] + (size_t)r0 * src_pitch[k], dst_pitch[k], src_pitch[k], row_bytes[k],
                                  r0 + per < rows[k] ? per : rows[k] - r0};
            p->chunk[n++] = j;
        }
    }
    __atomic_store_n(&p->done, 0, __ATOMIC_RELEASE);
    const uint64_t pic = ((__atomic_load_n(&p->ticket, __ATOMIC_RELAXED) >> 32) + 1) & 0xffffffffu;
    pthread_mutex_lock(&p->mu);
    __atomic_store_n(&p->ticket, pic << 32 | (uint64_t)n << 16, __ATOMIC_RELEASE);
    pthread_cond_broadcast(&p->cv_go);
    pthread_mutex_unlock(&p->mu);
    copy_claim(p, pic);
    for (int i = 0; i < sws_spin && __atomic_load_n(&p->done, __ATOMIC_ACQUIRE) < n; i++) cpu_relax();
    if (__atomic_load_n(&p->done, __ATOMIC_ACQUIRE) < n) {
        pthread_mutex_lock(&p->mu);
        while (__atomic_load_n(&p->done, __ATOMIC_ACQUIRE) < n) pthread_cond_wait(&p->cv_done, &p->mu);
        pthread_mutex_unlock(&p->mu);
    }
}

void *b2_sws_rt_create(int w, int h, int fmt);
void b2_sws_rt_free(void *rt);
int b2_sws_rt_scale(void *rt, const uint8_t *const src[], const int srcStride[], uint8_t *const dst[], const int dstStride[]);

struct b2_sws_context {
    int w, h, fmt, host_output;
    size_t in_bytes;
    void *rt;                   /* GPU round-trip state (device, buffers, stream) */
    copy_pool_t *pool;          /* helper threads of the staging copy (created with the first large picture) */
    int stats;                  /* B2ENC_STATS=1: seconds spent waiting for a staging buffer's upload / copying, printed by freeContext */
    double st_wait, st_copy;
    long st_pictures;
};
static double sws_now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

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
    { const char *ev = getenv("B2ENC_STATS"); c->stats = ev && atoi(ev) > 0; }
    return c;
}

void b2_sws_freeContext(b2_sws_context_t *c)
{
    if (!c) return;
    if (c->stats && c->st_pictures)
        fprintf(stderr, "b2enc stats: sws_scale into the encoder picture: %ld pictures, %.3f s waiting for a staging buffer's upload, %.3f s copying "
                        "(%.0f us per picture, caller + %d helpers)\n", c->st_pictures, c->st_wait, c->st_copy, 1e6 * c->st_copy / (double)c->st_pictures,
                c->pool ? c->pool->n : 0);
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
        const double s0 = c->stats ? sws_now() : 0.0;
        uint8_t *p = b2h_picture_stage(rec, c->in_bytes);
        const double s1 = c->stats ? sws_now() : 0.0;
        if (!p) { fprintf(stderr, "b2enc: b2_sws_scale: cannot allocate page-locked staging\n"); return -1; }
        if (!c->pool && c->in_bytes >= ((size_t)2 << 20)) {
            long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
            const char *e = getenv("B2ENC_SWS_SPIN");
            if (e && atoi(e) >= 0) sws_spin = atoi(e);
            e = getenv("B2ENC_SWS_THREADS");
            int helpers = e ? atoi(e) - 1 : 3;
            if (ncpu > 0 && helpers > ncpu / 4) helpers = (int)(ncpu / 4);
            c->pool = copy_pool_create(helpers < 0 ? 0 : helpers);
        }
        int rb[3], rws[3];
        const int np = b2_fmt_layout(c->fmt, w, h, rb, rws);
        uint8_t *dp[3]; size_t dpitch[3], spitch[3], rbytes[3];
        for (int k = 0; k < np; k++) {
            dp[k] = p; dpitch[k] = rbytes[k] = (size_t)rb[k]; spitch[k] = (size_t)srcStride[k];
            p += (size_t)rb[k] * rws[k];
        }
        copy_planes(c->pool, np, dp, dpitch, src, spitch, rbytes, rws);
        if (c->stats) { c->st_wait += s1 - s0; c->st_copy += sws_now() - s1; c->st_pictures++; }
        rec->fmt = c->fmt;
        rec->deferred = 1;
        return h;
    }
    if (rec) rec->deferred = 0;                           /* its planes are about to hold a converted frame */
    return b2_sws_rt_scale(c->rt, src, srcStride, dst, dstStride);
}
