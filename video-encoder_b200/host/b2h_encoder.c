/*
 * b2h_encoder.c -- host side of the drop-in encoder boundary (include/b2enc.h): the x264 API subset
 * that av_encode.c calls (enc_x264_open :378-438, the encode loop :968-975, the drain loop :1076-1083),
 * implemented over the CUDA engine (include/b2enc_engine.h) and the host entropy stage (b2h_entropy.h: CABAC or CAVLC).
 *
 * Frame queue (row a0 of SURVEY.md 8a): the caller still hands over one picture per call, but pictures
 * are gathered into i_gop_slots closed GOPs of i_keyint_max frames; a full batch is advanced through
 * the GPU in lock-step (one launch sequence per frame index covers all GOPs), the host entropy-codes
 * the per-MB results and frames are then returned in display order, one per call -- exactly the
 * "0 = no output yet / delayed_frames() / encode(NULL) drains" contract main() relies on
 * (av_encode.c:971-974, :1076-1083).  A pipeline thread drives batch k through the GPU and the entropy workers while the
 * caller already gathers batch k+1 into the other half of the pinned input ring, so copying pictures in and encoding
 * overlap.  i_gop_slots = 1 encodes every picture immediately (zero delay).
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "b2enc.h"
#include "b2enc_engine.h"
#include "b2h_entropy.h"

void *b2_pinned_alloc(size_t n);
void b2_pinned_free(void *p);

typedef struct {
    uint8_t *data;
    int size;
    int nal_count;
    int nal_off[3], nal_size[3], nal_type[3], nal_ref[3];
    int64_t pts;
    int key;
} outframe_t;

#define B2_MAX_WORKERS 16

/* one frame to entropy-code; the per-MB decisions and the packed levels are heap copies of the engine's pinned result set, so
 * the GPU may run ahead of the entropy workers by any number of steps (the engine keeps only two result sets) */
typedef struct { int qi, slot, t; int64_t gop_index; uint8_t *res; size_t packed_bytes; } job_t;

struct b2_encoder {
    b2_param_t p;
    int qp, S, L, mbw, mbh, nmb;
    b2_engine_t *eng;
    b2h_entropy_t *ent;
    b2h_seq_t seq;
    /* entropy worker pool: one frame per job, queued per batch in GPU completion order; each worker owns its neighbour-map
     * scratch.  Frames of one GOP are independent for the entropy stage (contexts reset per slice), so frame t+1 of a GOP
     * may be coded while frame t still is -- all workers stay busy even with fewer GOP slots than host cores. */
    int nworkers;
    pthread_t workers[B2_MAX_WORKERS];
    b2h_entropy_t *went[B2_MAX_WORKERS];
    uint8_t *wscratch[B2_MAX_WORKERS];
    pthread_mutex_t mu;
    pthread_cond_t cv_job, cv_done;
    job_t *jobs;
    int njobs, next_job, done_jobs, job_error, stop;
    int64_t *pts;              /* [2][S*L] pts of the frames of the two batch halves */
    int batch_frames;          /* frames gathered into the current half */
    /* asynchronous batch pipeline (S > 1): the caller gathers batch k+1 into one half of the input ring while the pipeline
     * thread drives batch k (GPU steps + entropy workers) out of the other half */
    pthread_t pipe_thread;
    int pipe_started, pipe_busy, pipe_stop, pipe_error;
    pthread_mutex_t pmu;
    pthread_cond_t pcv_submit, pcv_done;
    int gather_half, sub_half, sub_frames;
    int inflight;              /* frames handed to the pipeline thread that are not in the fifo yet */
    outframe_t *fifo;          /* finished frames in display order */
    int fifo_cap, fifo_head, fifo_count;
    int gop_pos;               /* zero-delay mode: position inside the current GOP */
    outframe_t *outq;          /* [S*L] finished frames of the last processed batch, display order */
    int out_head, out_count;
    int64_t gops_done;
    uint8_t *scratch;
    size_t scratch_cap;
    uint8_t *ret_buf;          /* payload of the frame returned by the last call */
    b2_nal_t nals[3];
};

static const struct { const char *name; int merange, subpel, intra_in_p; } presets[] = {
    {"ultrafast", 16, 0, 0}, {"superfast", 16, 1, 1}, {"veryfast", 16, 1, 1}, {"faster", 16, 1, 1}, {"fast", 16, 1, 1},
    {"medium", 16, 1, 1},    {"slow", 32, 1, 1},      {"slower", 32, 1, 1},   {"veryslow", 32, 1, 1}, {"placebo", 32, 1, 1}};
static const char *tunes[] = {"film", "animation", "grain", "stillimage", "psnr", "ssim", "fastdecode", "zerolatency"};

int b2_param_default_preset(b2_param_t *p, const char *preset, const char *tune)
{
    if (!p) return -1;
    memset(p, 0, sizeof(*p));
    p->i_fps_num = 25; p->i_fps_den = 1;
    p->vui.i_sar_width = p->vui.i_sar_height = 0;
    p->rc.i_rc_method = B2_RC_CRF; p->rc.f_rf_constant = 23.0f; p->rc.i_qp_constant = 26;
    p->b_annexb = 1;
    /* 16 closed GOPs in lock-step: a step of only 8 frames leaves the GPU waiting on the per-frame latency chain (K7 / K8) */
    p->i_keyint_max = 32; p->i_gop_slots = 16; p->i_device = 0; p->i_csp_in = B2_FMT_YUV420P;
    p->b_deblocking_filter = 1;
    p->b_cabac = 1;                      /* x264 default at every preset but ultrafast */
    p->b_transform_8x8 = 0;
    int found = preset == NULL;
    p->i_merange = 16; p->b_subpel = 1; p->b_intra_in_p = 1;
    for (unsigned i = 0; preset && i < sizeof(presets) / sizeof(presets[0]); i++)
        if (!strcmp(preset, presets[i].name)) {
            p->i_merange = presets[i].merange; p->b_subpel = presets[i].subpel; p->b_intra_in_p = presets[i].intra_in_p;
            found = 1;
        }
    if (!found) { fprintf(stderr, "b2enc: invalid preset '%s'\n", preset); return -1; }
    if (preset && !strcmp(preset, "ultrafast")) { p->b_cabac = 0; p->b_deblocking_filter = 0; }     /* as x264's ultrafast */
    if (tune) {
        int ok = 0;
        for (unsigned i = 0; i < sizeof(tunes) / sizeof(tunes[0]); i++) ok |= !strcmp(tune, tunes[i]);
        if (!ok) { fprintf(stderr, "b2enc: invalid tune '%s'\n", tune); return -1; }
        if (!strcmp(tune, "zerolatency")) p->i_gop_slots = 1;
    }
    return 0;
}

int b2_param_apply_profile(b2_param_t *p, const char *profile)
{
    if (!p) return -1;
    if (!profile) return 0;
    /* like x264_param_apply_profile: a profile only removes tools.  baseline: CAVLC, 4x4 transform; main: no 8x8 transform */
    if (!strcmp(profile, "baseline")) { p->b_cabac = 0; p->b_transform_8x8 = 0; return 0; }
    if (!strcmp(profile, "main")) { p->b_transform_8x8 = 0; return 0; }
    if (!strcmp(profile, "high")) return 0;
    fprintf(stderr, "b2enc: invalid profile: %s\n", profile);
    return -1;
}

int b2_picture_alloc(b2_picture_t *pic, int i_csp, int i_width, int i_height)
{
    if (!pic || i_csp != B2_CSP_I420 || i_width < 2 || i_height < 2) return -1;
    memset(pic, 0, sizeof(*pic));
    int cw = (i_width + 1) / 2, ch = (i_height + 1) / 2;
    size_t ny = (size_t)i_width * i_height, nc = (size_t)cw * ch;
    uint8_t *buf = (uint8_t *)b2_pinned_alloc(ny + 2 * nc);          /* pinned: H2D copies run at full PCIe rate */
    if (!buf) return -1;
    pic->img.i_csp = i_csp; pic->img.i_plane = 3;
    pic->img.plane[0] = buf; pic->img.plane[1] = buf + ny; pic->img.plane[2] = buf + ny + nc;
    pic->img.i_stride[0] = i_width; pic->img.i_stride[1] = cw; pic->img.i_stride[2] = cw;
    pic->opaque = buf;
    return 0;
}

void b2_picture_clean(b2_picture_t *pic)
{
    if (!pic) return;
    b2_pinned_free(pic->opaque);
    memset(pic, 0, sizeof(*pic));
}

static void *worker_main(void *arg);
static void *pipe_main(void *arg);

b2_t *b2_encoder_open(b2_param_t *p)
{
    if (!p || p->i_width < 16 || p->i_height < 16) { fprintf(stderr, "b2enc: bad picture size\n"); return NULL; }
    b2_t *h = (b2_t *)calloc(1, sizeof(*h));
    if (!h) return NULL;
    h->p = *p;
    int qp = p->rc.i_rc_method == B2_RC_CQP ? p->rc.i_qp_constant : (int)(p->rc.f_rf_constant + 0.5f);
    h->qp = qp < 10 ? 10 : (qp > 51 ? 51 : qp);      /* CRF is mapped to a constant QP (north_star: fixed QP) */
    h->S = p->i_gop_slots > 0 ? p->i_gop_slots : 1;
    h->L = p->i_keyint_max > 0 ? p->i_keyint_max : 32;
    b2_engine_cfg_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = p->i_device; cfg.width = p->i_width; cfg.height = p->i_height; cfg.slots = h->S;
    cfg.in_fmt = p->i_csp_in; cfg.in_ring = h->S == 1 ? 1 : 2 * h->L; cfg.merange = p->i_merange ? p->i_merange : 16; cfg.qp = h->qp;
    cfg.subpel = p->b_subpel; cfg.intra_in_p = p->b_intra_in_p; cfg.profile = 0; cfg.deblock = p->b_deblocking_filter;
    cfg.transform8x8 = p->b_transform_8x8 != 0;
    cfg.partitions = p->b_partitions < 0 ? 0 : (p->b_partitions > 2 ? 2 : p->b_partitions);
    cfg.pack_levels = 1;                              /* only blocks with non-zero levels cross PCIe (K9) */
    h->eng = b2_engine_create(&cfg);
    if (!h->eng) { free(h); return NULL; }
    int w16, h16;
    b2_engine_geometry(h->eng, &h->mbw, &h->mbh, &w16, &h16);
    h->nmb = h->mbw * h->mbh;
    h->ent = b2h_entropy_create(h->mbw, h->mbh);
    h->seq.width = p->i_width; h->seq.height = p->i_height; h->seq.fps_num = p->i_fps_num; h->seq.fps_den = p->i_fps_den;
    h->seq.sar_w = p->vui.i_sar_width; h->seq.sar_h = p->vui.i_sar_height; h->seq.qp = h->qp;
    h->seq.deblock = p->b_deblocking_filter;
    h->seq.cabac = p->b_cabac != 0; h->seq.transform8x8 = p->b_transform_8x8 != 0;
    h->pts = (int64_t *)calloc((size_t)2 * h->S * h->L, sizeof(int64_t));
    h->fifo_cap = 3 * h->S * h->L;
    h->fifo = (outframe_t *)calloc((size_t)h->fifo_cap, sizeof(outframe_t));
    pthread_mutex_init(&h->pmu, NULL); pthread_cond_init(&h->pcv_submit, NULL); pthread_cond_init(&h->pcv_done, NULL);
    h->outq = (outframe_t *)calloc((size_t)h->S * h->L, sizeof(outframe_t));
    h->scratch_cap = (size_t)h->nmb * 3072 + 65536;
    h->scratch = (uint8_t *)malloc(h->scratch_cap);
    if (!h->ent || !h->pts || !h->outq || !h->scratch || !h->fifo) { b2_encoder_close(h); return NULL; }
    pthread_mutex_init(&h->mu, NULL); pthread_cond_init(&h->cv_job, NULL); pthread_cond_init(&h->cv_done, NULL);
    if (h->S > 1) {
        long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
        int nw = B2_MAX_WORKERS;                         /* frames, not GOPs, are the unit of work: use the cores there are */
        if (ncpu > 0 && nw > ncpu) nw = (int)ncpu;
        h->jobs = (job_t *)calloc((size_t)h->S * h->L, sizeof(job_t));
        for (int i = 0; i < nw; i++) {
            h->went[i] = b2h_entropy_create(h->mbw, h->mbh);
            h->wscratch[i] = (uint8_t *)malloc(h->scratch_cap);
            if (!h->jobs || !h->went[i] || !h->wscratch[i]) { b2_encoder_close(h); return NULL; }
        }
        pthread_mutex_lock(&h->mu);                      /* workers look themselves up in h->workers[] under the lock */
        for (int i = 0; i < nw; i++) {
            if (pthread_create(&h->workers[i], NULL, worker_main, h)) break;
            h->nworkers++;
        }
        pthread_mutex_unlock(&h->mu);
        if (pthread_create(&h->pipe_thread, NULL, pipe_main, h)) { b2_encoder_close(h); return NULL; }
        h->pipe_started = 1;
    }
    return h;
}

void b2_encoder_close(b2_t *h)
{
    if (!h) return;
    if (h->pipe_started) {
        pthread_mutex_lock(&h->pmu);
        h->pipe_stop = 1;
        pthread_cond_broadcast(&h->pcv_submit);
        pthread_mutex_unlock(&h->pmu);
        pthread_join(h->pipe_thread, NULL);
    }
    if (h->fifo)
        for (int i = 0; i < h->fifo_cap; i++) free(h->fifo[i].data);
    free(h->fifo);
    if (h->nworkers > 0) {
        pthread_mutex_lock(&h->mu);
        h->stop = 1;
        pthread_cond_broadcast(&h->cv_job);
        pthread_mutex_unlock(&h->mu);
        for (int i = 0; i < h->nworkers; i++) pthread_join(h->workers[i], NULL);
    }
    for (int i = 0; i < B2_MAX_WORKERS; i++) { b2h_entropy_destroy(h->went[i]); free(h->wscratch[i]); }
    if (h->jobs)
        for (int i = h->next_job; i < h->njobs; i++) free(h->jobs[i].res);          /* frames queued but never coded (error paths) */
    free(h->jobs);
    if (h->outq)
        for (int i = 0; i < h->S * h->L; i++) free(h->outq[i].data);
    free(h->outq); free(h->pts); free(h->scratch); free(h->ret_buf);
    b2h_entropy_destroy(h->ent);
    b2_engine_destroy(h->eng);
    free(h);
}

/* picture -> engine input ring: DMA from the page-locked picture when it is one (no host copy), else through the pinned staging */
static int put_picture(b2_t *h, int slot, int ring, const b2_picture_t *pic)
{
    int rc = b2_engine_put_frame_direct(h->eng, slot, ring, (const uint8_t *const *)pic->img.plane, pic->img.i_stride);
    if (rc > 0) rc = b2_engine_put_frame(h->eng, slot, ring, (const uint8_t *const *)pic->img.plane, pic->img.i_stride);
    return rc;
}

static void put_prefix(uint8_t *d, int annexb, size_t nal_size)
{
    if (annexb) { d[0] = 0; d[1] = 0; d[2] = 0; d[3] = 1; }
    else { d[0] = (uint8_t)(nal_size >> 24); d[1] = (uint8_t)(nal_size >> 16); d[2] = (uint8_t)(nal_size >> 8); d[3] = (uint8_t)nal_size; }
}

/* entropy-code one frame's results into outq[qi] */
static int finish_frame(b2_t *h, b2h_entropy_t *ent, uint8_t *s, const b2_mbinfo_t *info, const uint8_t *packed, size_t packed_bytes,
                        int qi, int t, int64_t gop_index)
{
    outframe_t *o = &h->outq[qi];
    size_t pos = 0;
    o->nal_count = 0;
    const int is_idr = t == 0;
    if (is_idr) {
        for (int k = 0; k < 2; k++) {
            size_t n = k == 0 ? b2h_write_sps(&h->seq, s + pos + 4, h->scratch_cap - pos - 4)
                              : b2h_write_pps(&h->seq, s + pos + 4, h->scratch_cap - pos - 4);
            if (!n) return -1;
            put_prefix(s + pos, h->p.b_annexb, n);
            o->nal_off[o->nal_count] = (int)pos; o->nal_size[o->nal_count] = (int)n + 4;
            o->nal_type[o->nal_count] = k == 0 ? B2_NAL_SPS : B2_NAL_PPS; o->nal_ref[o->nal_count] = 3;
            o->nal_count++;
            pos += n + 4;
        }
    }
    if (!info || !packed) return -1;
    size_t n = b2h_write_slice_packed(ent, &h->seq, is_idr ? B2_FRAME_I : B2_FRAME_P, t, (int)(gop_index & 0xffff), info, packed,
                                      packed_bytes, s + pos + 4, h->scratch_cap - pos - 4);
    if (!n) { fprintf(stderr, "b2enc: slice buffer overflow\n"); return -1; }
    put_prefix(s + pos, h->p.b_annexb, n);
    o->nal_off[o->nal_count] = (int)pos; o->nal_size[o->nal_count] = (int)n + 4;
    o->nal_type[o->nal_count] = is_idr ? B2_NAL_SLICE_IDR : B2_NAL_SLICE; o->nal_ref[o->nal_count] = is_idr ? 3 : 2;
    o->nal_count++;
    pos += n + 4;
    free(o->data);
    o->data = (uint8_t *)malloc(pos);
    if (!o->data) return -1;
    memcpy(o->data, s, pos);
    o->size = (int)pos;
    o->key = is_idr;
    return 0;
}

static void *worker_main(void *arg)
{
    b2_t *h = (b2_t *)arg;
    int me = -1;
    pthread_mutex_lock(&h->mu);
    for (int i = 0; i < h->nworkers; i++)
        if (pthread_equal(h->workers[i], pthread_self())) me = i;
    for (;;) {
        while (!h->stop && h->next_job >= h->njobs) pthread_cond_wait(&h->cv_job, &h->mu);
        if (h->stop) break;
        job_t j = h->jobs[h->next_job++];
        pthread_mutex_unlock(&h->mu);
        int rc = finish_frame(h, h->went[me], h->wscratch[me], (const b2_mbinfo_t *)j.res, j.res + (size_t)h->nmb * sizeof(b2_mbinfo_t),
                              j.packed_bytes, j.qi, j.t, j.gop_index);
        free(j.res);
        pthread_mutex_lock(&h->mu);
        if (rc) h->job_error = 1;
        h->done_jobs++;
        pthread_cond_broadcast(&h->cv_done);
    }
    pthread_mutex_unlock(&h->mu);
    return NULL;
}

/* queue frame t of GOPs [0,nt) for the worker pool (or code it inline when there is none): the result set behind `ticket`
 * is copied out, so it may be overwritten as soon as this returns */
static int entropy_step(b2_t *h, int t, int nt, int ticket, const int64_t *pts)
{
    for (int g = 0; g < nt; g++) {
        const int qi = g * h->L + t;
        const b2_mbinfo_t *info = b2_engine_info_ticket(h->eng, ticket, g);
        size_t packed_bytes = 0;
        const uint8_t *packed = b2_engine_packed_ticket(h->eng, ticket, g, &packed_bytes);
        if (!info || !packed) return -1;
        h->outq[qi].pts = pts[qi];
        if (h->nworkers == 0) {
            if (finish_frame(h, h->ent, h->scratch, info, packed, packed_bytes, qi, t, h->gops_done + g)) return -1;
            continue;
        }
        const size_t ni = (size_t)h->nmb * sizeof(b2_mbinfo_t);
        job_t j = {qi, g, t, h->gops_done + g, (uint8_t *)malloc(ni + packed_bytes + 1), packed_bytes};
        if (!j.res) return -1;
        memcpy(j.res, info, ni);
        memcpy(j.res + ni, packed, packed_bytes);
        pthread_mutex_lock(&h->mu);
        h->jobs[h->njobs++] = j;
        pthread_cond_signal(&h->cv_job);
        pthread_mutex_unlock(&h->mu);
    }
    return 0;
}

/* wait until every queued frame of the batch is coded, then reset the queue */
static int entropy_drain(b2_t *h)
{
    if (h->nworkers == 0) return 0;
    pthread_mutex_lock(&h->mu);
    while (h->done_jobs < h->njobs) pthread_cond_wait(&h->cv_done, &h->mu);
    h->njobs = 0; h->next_job = 0; h->done_jobs = 0;
    const int err = h->job_error;
    pthread_mutex_unlock(&h->mu);
    return err ? -1 : 0;
}

static int issue_step(b2_t *h, int half, int t, int nt)
{
    const int ring = h->S == 1 ? 0 : half * h->L + t;
    if (b2_engine_h2d(h->eng, 0, nt, ring)) return -1;
    if (b2_engine_encode(h->eng, t == 0 ? B2_FRAME_I : B2_FRAME_P, nt, ring)) return -1;
    return b2_engine_d2h(h->eng, nt);
}

/* advance one gathered batch (possibly partial) through the GPU and the entropy stage: while the host entropy-codes
 * frame t of every GOP, the GPU already encodes frame t+1 (result sets are double buffered).  Runs on the pipeline thread. */
static int process_batch(b2_t *h, int half, int n)
{
    const int L = h->L;
    const int ngop = (n + L - 1) / L, last_len = n - (ngop - 1) * L;
    const int64_t *pts = h->pts + (size_t)half * h->S * L;
    int nt = ngop;                                         /* GOPs that have a frame 0 */
    if (issue_step(h, half, 0, nt)) return -1;
    for (int t = 0; t < L && nt > 0; t++) {
        const int ticket = b2_engine_ticket(h->eng);
        const int nt_next = t + 1 < L ? (t + 1 < last_len ? ngop : ngop - 1) : 0;     /* a prefix of the slots */
        if (b2_engine_wait_ticket(h->eng, ticket)) return -1;
        if (nt_next > 0 && issue_step(h, half, t + 1, nt_next)) return -1;
        if (entropy_step(h, t, nt, ticket, pts)) return -1;
        nt = nt_next;
    }
    if (entropy_drain(h)) return -1;
    if (b2_engine_sync(h->eng)) return -1;
    h->gops_done += ngop;
    pthread_mutex_lock(&h->pmu);                          /* hand the frames over in display order */
    for (int i = 0; i < n; i++) {
        outframe_t *dst = &h->fifo[(h->fifo_head + h->fifo_count) % h->fifo_cap];
        free(dst->data);
        *dst = h->outq[i];
        h->outq[i].data = NULL;
        h->fifo_count++;
    }
    h->inflight -= n;
    pthread_mutex_unlock(&h->pmu);
    return 0;
}

static void *pipe_main(void *arg)
{
    b2_t *h = (b2_t *)arg;
    pthread_mutex_lock(&h->pmu);
    for (;;) {
        while (!h->pipe_busy && !h->pipe_stop) pthread_cond_wait(&h->pcv_submit, &h->pmu);
        if (h->pipe_stop) break;
        const int half = h->sub_half, n = h->sub_frames;
        pthread_mutex_unlock(&h->pmu);
        const int rc = process_batch(h, half, n);
        pthread_mutex_lock(&h->pmu);
        if (rc) h->pipe_error = 1;
        h->pipe_busy = 0;
        pthread_cond_broadcast(&h->pcv_done);
    }
    pthread_mutex_unlock(&h->pmu);
    return NULL;
}

/* hand the gathered half to the pipeline thread (waits until it has finished the previous batch) and gather into the other */
static void submit_batch(b2_t *h)
{
    pthread_mutex_lock(&h->pmu);
    while (h->pipe_busy) pthread_cond_wait(&h->pcv_done, &h->pmu);
    h->sub_half = h->gather_half; h->sub_frames = h->batch_frames;
    h->inflight += h->batch_frames;
    h->pipe_busy = 1;
    pthread_cond_signal(&h->pcv_submit);
    pthread_mutex_unlock(&h->pmu);
    h->gather_half ^= 1;
    h->batch_frames = 0;
}

int b2_encoder_delayed_frames(b2_t *h)
{
    if (!h) return 0;
    if (h->S == 1) return h->out_count;
    pthread_mutex_lock(&h->pmu);
    const int n = h->batch_frames + h->inflight + h->fifo_count;
    pthread_mutex_unlock(&h->pmu);
    return n;
}

static int return_frame(b2_t *h, outframe_t *o, b2_nal_t **pp_nal, int *pi_nal, b2_picture_t *pic_out)
{
    free(h->ret_buf);
    h->ret_buf = o->data; o->data = NULL;                           /* hand the payload over; valid until the next call */
    for (int i = 0; i < o->nal_count; i++) {
        h->nals[i].i_ref_idc = o->nal_ref[i]; h->nals[i].i_type = o->nal_type[i];
        h->nals[i].i_payload = o->nal_size[i]; h->nals[i].p_payload = h->ret_buf + o->nal_off[i];
    }
    *pp_nal = h->nals; *pi_nal = o->nal_count;
    if (pic_out) {
        memset(pic_out, 0, sizeof(*pic_out));
        pic_out->i_pts = o->pts; pic_out->i_dts = o->pts; pic_out->b_keyframe = o->key;
        pic_out->i_type = o->key ? B2_TYPE_IDR : B2_TYPE_P;
    }
    return o->size;
}

int b2_encoder_encode(b2_t *h, b2_nal_t **pp_nal, int *pi_nal, b2_picture_t *pic_in, b2_picture_t *pic_out)
{
    if (!h || !pp_nal || !pi_nal) return -1;
    *pi_nal = 0; *pp_nal = NULL;
    if (h->S == 1) {
        /* zero-delay mode: one slot, encode every picture as it arrives */
        if (!pic_in) return 0;
        const int t = h->gop_pos;
        if (put_picture(h, 0, 0, pic_in)) return -1;
        if (b2_engine_h2d(h->eng, 0, 1, 0) || b2_engine_encode(h->eng, t == 0 ? B2_FRAME_I : B2_FRAME_P, 1, 0) ||
            b2_engine_d2h(h->eng, 1) || b2_engine_sync(h->eng))
            return -1;
        size_t packed_bytes = 0;
        const uint8_t *packed = b2_engine_packed(h->eng, 0, &packed_bytes);
        if (finish_frame(h, h->ent, h->scratch, b2_engine_info(h->eng, 0), packed, packed_bytes, 0, t, h->gops_done)) return -1;
        h->outq[0].pts = pic_in->i_pts;
        h->gop_pos = t + 1;
        if (h->gop_pos == h->L) { h->gop_pos = 0; h->gops_done++; }
        return return_frame(h, &h->outq[0], pp_nal, pi_nal, pic_out);
    }
    if (pic_in) {
        const int idx = h->batch_frames, g = idx / h->L, t = idx % h->L;
        if (put_picture(h, g, h->gather_half * h->L + t, pic_in)) return -1;
        h->pts[(size_t)h->gather_half * h->S * h->L + idx] = pic_in->i_pts;
        h->batch_frames++;
        if (h->batch_frames == h->S * h->L) submit_batch(h);        /* full: the pipeline thread takes it, gathering goes on */
    } else {
        if (h->batch_frames > 0) submit_batch(h);                   /* flush: partial batch */
        pthread_mutex_lock(&h->pmu);                                /* a flush call waits for its frame (av_encode.c:1076-1083) */
        while (h->fifo_count == 0 && h->pipe_busy && !h->pipe_error) pthread_cond_wait(&h->pcv_done, &h->pmu);
        pthread_mutex_unlock(&h->pmu);
    }
    pthread_mutex_lock(&h->pmu);
    if (h->pipe_error) { pthread_mutex_unlock(&h->pmu); fprintf(stderr, "b2enc: encode pipeline failed\n"); return -1; }
    if (h->fifo_count == 0) { pthread_mutex_unlock(&h->pmu); return 0; }
    outframe_t o = h->fifo[h->fifo_head];
    h->fifo[h->fifo_head].data = NULL;
    h->fifo_head = (h->fifo_head + 1) % h->fifo_cap;
    h->fifo_count--;
    pthread_mutex_unlock(&h->pmu);
    return return_frame(h, &o, pp_nal, pi_nal, pic_out);
}
