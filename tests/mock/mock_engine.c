/*
 * mock_engine.c -- TEST INFRASTRUCTURE ONLY (never linked into libb2enc.so).
 *
 * A CPU stand-in for the CUDA engine behind include/b2enc_engine.h, built on the oracle (oracle/b2o_*.c), so that the
 * HOST logic of the drop-in encoder (video-encoder_b200/host/b2h_encoder.c: GOP streaming, one thread per GPU, entropy
 * worker pool, display-order fifo, flush / delay contract, deferred sws_scale, format switches) can be exercised by the
 * `-m "not gpu"` suite on a machine without a GPU.  tests/test_host_pipeline.py links this file with the product's host
 * sources into tests/mock/_build/libb2enc_mock.so.  It implements only the engine calls b2h_encoder.c makes; results
 * "land" asynchronously (b2_engine_group_done says "not yet" the first time it is asked) to exercise the polling paths.
 * B2_MOCK_DEVICES = number of pretend GPUs.
 *
 * B2_MOCK_CANNED=<dir> (scripts/host_ceiling.py): a ZERO-LATENCY engine for measuring what the host stage alone can carry.
 * The directory holds oracle results of one I and a few P frames at the encoder's size (info<k>.bin = b2_mbinfo_t[], packed<k>.bin
 * = packed levels; k = 0 is the I frame); every step "lands" at once with the next canned result, pictures are not copied
 * (on the GPU the DMA engine does that), so all that runs is the product's host code on realistic decisions and levels.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "b2enc_engine.h"
#include "b2enc_kernels.h"
#include "../../oracle/b2o.h"

typedef struct {
    uint8_t *in;                    /* [ring][in_bytes] raw pictures */
    b2o_frame_t cur, rec[2];
    int ref_idx;
    b2_mv_t *prev_mv;
    b2_mbinfo_t *info[2];           /* result sets */
    b2_mbinfo_packed_t *pinfo[2];   /* the same as the 24-byte records that cross PCIe (K9a on the GPU) */
    uint8_t *packed[2];
    size_t packed_bytes[2];
    int res_set, host_set, polled[2];
    b2_mbcoef_t *coef;
} mslot_t;

#define CANNED_MAX 8
typedef struct { b2_mbinfo_t *info; b2_mbinfo_packed_t *pinfo; uint8_t *packed; size_t packed_bytes; } canned_t;

struct b2_engine {
    canned_t canned[CANNED_MAX];
    int ncanned;                    /* > 0: zero-latency mode */
    long *canned_step;              /* per slot: P frames served */
    b2_engine_cfg_t cfg;
    int mbw, mbh, nmb, w16, h16;
    size_t in_bytes;
    mslot_t *slots;
    long launches, puts;
};

static size_t input_bytes(int fmt, int w, int h)
{
    int rb[3], rows[3];
    const int np = b2_fmt_layout(fmt, w, h, rb, rows);
    size_t n = 0;
    for (int p = 0; p < np; p++) n += (size_t)rb[p] * rows[p];
    return np && b2_fmt_size_ok(fmt, w, h) ? n : 0;
}

int b2_device_count(void) { const char *e = getenv("B2_MOCK_DEVICES"); return e ? atoi(e) : 1; }
int b2_device_mem_info(int device, size_t *f, size_t *t) { (void)device; if (f) *f = (size_t)64 << 30; if (t) *t = (size_t)64 << 30; return 0; }
void *b2_pinned_alloc(size_t n) { return malloc(n); }
void b2_pinned_free(void *p) { free(p); }

/* round-trip conversion of b2_sws_scale (csrc/b2_sws.cu in the product): the oracle's closed forms */
typedef struct { int w, h, fmt; } mrt_t;
void *b2_sws_rt_create(int w, int h, int fmt) { if (b2_device_count() <= 0) return NULL; mrt_t *r = calloc(1, sizeof(*r)); r->w = w; r->h = h; r->fmt = fmt; return r; }
void b2_sws_rt_free(void *rt) { free(rt); }
int b2_sws_rt_scale(void *rt, const uint8_t *const src[], const int srcStride[], uint8_t *const dst[], const int dstStride[])
{
    mrt_t *r = (mrt_t *)rt;
    const uint8_t *s4[4] = {src[0], src[1], src[2], NULL};
    int st4[4] = {srcStride[0], srcStride[1], srcStride[2], 0};
    return b2o_convert_to_i420(r->fmt, r->w, r->h, s4, st4, dst, dstStride) ? -1 : r->h;
}

b2_engine_t *b2_engine_create(const b2_engine_cfg_t *cfg)
{
    if (b2_device_count() <= cfg->device) return NULL;
    b2_engine_t *e = calloc(1, sizeof(*e));
    e->cfg = *cfg;
    e->w16 = (cfg->width + 15) & ~15; e->h16 = (cfg->height + 15) & ~15;
    e->mbw = e->w16 / 16; e->mbh = e->h16 / 16; e->nmb = e->mbw * e->mbh;
    e->in_bytes = input_bytes(cfg->in_fmt, cfg->width, cfg->height);
    if (!e->in_bytes || cfg->streams != cfg->slots) { free(e); return NULL; }          /* the drop-in uses one group per slot */
    e->slots = calloc((size_t)cfg->slots, sizeof(mslot_t));
    const char *cdir = getenv("B2_MOCK_CANNED");
    for (int k = 0; cdir && cdir[0] && k < CANNED_MAX; k++) {
        char pi[600], pp[600];
        snprintf(pi, sizeof pi, "%s/info%d.bin", cdir, k); snprintf(pp, sizeof pp, "%s/packed%d.bin", cdir, k);
        FILE *fi = fopen(pi, "rb"), *fp = fopen(pp, "rb");
        if (!fi || !fp) { if (fi) fclose(fi); if (fp) fclose(fp); break; }
        canned_t *c = &e->canned[k];
        c->info = malloc((size_t)e->nmb * sizeof(b2_mbinfo_t)); c->pinfo = malloc((size_t)e->nmb * sizeof(b2_mbinfo_packed_t));
        fseek(fp, 0, SEEK_END); c->packed_bytes = (size_t)ftell(fp); fseek(fp, 0, SEEK_SET);
        c->packed = malloc(c->packed_bytes + 1);
        const int ok = fread(c->info, sizeof(b2_mbinfo_t), (size_t)e->nmb, fi) == (size_t)e->nmb && fread(c->packed, 1, c->packed_bytes, fp) == c->packed_bytes;
        fclose(fi); fclose(fp);
        if (!ok) { fprintf(stderr, "mock engine: %s does not match %dx%d\n", pi, cfg->width, cfg->height); free(e->slots); free(e); return NULL; }
        for (int i = 0; i < e->nmb; i++) c->pinfo[i] = b2_mbinfo_pack(&c->info[i]);
        e->ncanned = k + 1;
    }
    if (e->ncanned == 1) e->ncanned = 0;                  /* needs an I and at least one P frame */
    e->canned_step = calloc((size_t)cfg->slots, sizeof(long));
    for (int s = 0; s < cfg->slots; s++) {
        mslot_t *m = &e->slots[s];
        if (!e->ncanned) {                                /* zero-latency mode touches none of the oracle state */
            m->in = malloc(e->in_bytes * cfg->in_ring);
            b2o_frame_alloc(&m->cur, cfg->width, cfg->height);
            b2o_frame_alloc(&m->rec[0], cfg->width, cfg->height); b2o_frame_alloc(&m->rec[1], cfg->width, cfg->height);
            m->prev_mv = calloc((size_t)e->nmb, sizeof(b2_mv_t));
            m->coef = calloc((size_t)e->nmb, sizeof(b2_mbcoef_t));
        }
        for (int k = 0; k < 2; k++) {
            m->info[k] = calloc((size_t)e->nmb, sizeof(b2_mbinfo_t));
            m->pinfo[k] = calloc((size_t)e->nmb, sizeof(b2_mbinfo_packed_t));
            m->packed[k] = malloc((size_t)e->nmb * sizeof(b2_mbcoef_t));
        }
    }
    return e;
}

void b2_engine_destroy(b2_engine_t *e)
{
    if (!e) return;
    for (int s = 0; s < e->cfg.slots; s++) {
        mslot_t *m = &e->slots[s];
        free(m->in); b2o_frame_free(&m->cur); b2o_frame_free(&m->rec[0]); b2o_frame_free(&m->rec[1]);
        free(m->prev_mv); free(m->coef);
        for (int k = 0; k < 2; k++) { free(m->info[k]); free(m->pinfo[k]); free(m->packed[k]); }
    }
    for (int k = 0; k < CANNED_MAX; k++) { free(e->canned[k].info); free(e->canned[k].pinfo); free(e->canned[k].packed); }
    free(e->canned_step);
    free(e->slots); free(e);
}

void b2_engine_geometry(const b2_engine_t *e, int *mbw, int *mbh, int *w16, int *h16)
{
    if (mbw) *mbw = e->mbw;
    if (mbh) *mbh = e->mbh;
    if (w16) *w16 = e->w16;
    if (h16) *h16 = e->h16;
}

int b2_engine_put_picture(b2_engine_t *e, int slot, int ring, const uint8_t *const src[4], const int stride[4])
{
    if (slot < 0 || slot >= e->cfg.slots || ring < 0 || ring >= e->cfg.in_ring) return -1;
    if (e->ncanned) return 0;                             /* the GPU's copy engine moves the picture: no host time */
    int rb[3], rows[3];
    const int np = b2_fmt_layout(e->cfg.in_fmt, e->cfg.width, e->cfg.height, rb, rows);
    uint8_t *d = e->slots[slot].in + (size_t)ring * e->in_bytes;
    for (int p = 0; p < np; p++) {
        for (int y = 0; y < rows[p]; y++) memcpy(d + (size_t)y * rb[p], src[p] + (size_t)y * stride[p], rb[p]);
        d += (size_t)rb[p] * rows[p];
    }
    return 0;
}

/* the mock reads the source at call time; the ticket only exercises the bookkeeping of the double-buffered staging */
long b2_engine_put_picture_async(b2_engine_t *e, int slot, int ring, const uint8_t *const src[4], const int stride[4])
{
    return b2_engine_put_picture(e, slot, ring, src, stride) ? -1 : ++e->puts;       /* caller thread only */
}
int b2_engine_put_wait(b2_engine_t *e, long ticket) { (void)e; (void)ticket; return 0; }

int b2_engine_set_input_format(b2_engine_t *e, int fmt)
{
    const size_t nb = input_bytes(fmt, e->cfg.width, e->cfg.height);
    if (!nb) return -1;
    for (int s = 0; s < e->cfg.slots && !e->ncanned; s++) { free(e->slots[s].in); e->slots[s].in = malloc(nb * e->cfg.in_ring); }
    e->cfg.in_fmt = fmt; e->in_bytes = nb;
    return 0;
}

int b2_engine_encode_group(b2_engine_t *e, int group, int frame_type, int ring)
{
    if (group < 0 || group >= e->cfg.slots || ring < 0 || ring >= e->cfg.in_ring) return -1;
    mslot_t *m = &e->slots[group];
    if (e->ncanned) {                                     /* the next canned result, into the other result set (as the GPU's copy-out does) */
        const int set = m->res_set ^ 1;
        const canned_t *c = &e->canned[frame_type == B2_FRAME_P ? 1 + (int)(e->canned_step[group]++ % (e->ncanned - 1)) : 0];
        memcpy(m->info[set], c->info, (size_t)e->nmb * sizeof(b2_mbinfo_t));
        memcpy(m->pinfo[set], c->pinfo, (size_t)e->nmb * sizeof(b2_mbinfo_packed_t));
        memcpy(m->packed[set], c->packed, c->packed_bytes);
        m->packed_bytes[set] = c->packed_bytes;
        m->res_set = set;
        e->launches++;
        return 0;
    }
    const int w = e->cfg.width, h = e->cfg.height, cw = (w + 1) / 2, ch = (h + 1) / 2;
    int rb[3], rows[3];
    const int np = b2_fmt_layout(e->cfg.in_fmt, w, h, rb, rows);
    const uint8_t *src[4] = {NULL, NULL, NULL, NULL}; int st[4] = {0, 0, 0, 0};
    const uint8_t *p = m->in + (size_t)ring * e->in_bytes;
    for (int k = 0; k < np; k++) { src[k] = p; st[k] = rb[k]; p += (size_t)rb[k] * rows[k]; }
    uint8_t *i420 = malloc((size_t)w * h + 2 * (size_t)cw * ch);
    uint8_t *dst[3] = {i420, i420 + (size_t)w * h, i420 + (size_t)w * h + (size_t)cw * ch};
    int dstr[3] = {w, cw, cw};
    if (b2o_convert_to_i420(e->cfg.in_fmt, w, h, src, st, dst, dstr)) { free(i420); return -1; }
    const uint8_t *cd[3] = {dst[0], dst[1], dst[2]};
    b2o_frame_load(&m->cur, cd, dstr);
    free(i420);
    b2o_params_t prm = {e->cfg.qp, e->cfg.merange, e->cfg.subpel, e->cfg.intra_in_p, e->cfg.deblock, e->cfg.transform8x8,
                        e->cfg.partitions, e->cfg.deblock_alpha, e->cfg.deblock_beta};
    const int set = m->res_set ^ 1;
    const int is_p = frame_type == B2_FRAME_P;
    memset(m->info[set], 0, (size_t)e->nmb * sizeof(b2_mbinfo_t));
    b2o_encode_frame(&prm, frame_type, &m->cur, is_p ? &m->rec[m->ref_idx] : NULL, &m->rec[m->ref_idx ^ 1], is_p ? m->prev_mv : NULL,
                     m->info[set], m->coef);
    for (int i = 0; i < e->nmb; i++) { m->prev_mv[i].x = m->info[set][i].mvx; m->prev_mv[i].y = m->info[set][i].mvy; }
    /* pack the present blocks (layout of include/b2enc_types.h, what K9 does on the GPU) */
    size_t pos = 0;
    for (int i = 0; i < e->nmb; i++) {
        uint32_t pm = b2_coef_present(&m->info[set][i]);
        for (int b = 0; b < B2_COEF_BLOCKS; b++)
            if (pm >> b & 1) { memcpy(m->packed[set] + pos, m->coef[i].blk[b], 32); pos += 32; }
    }
    m->packed_bytes[set] = pos;
    for (int i = 0; i < e->nmb; i++) m->pinfo[set][i] = b2_mbinfo_pack(&m->info[set][i]);
    m->ref_idx ^= 1; m->res_set = set;
    e->launches++;
    return 0;
}

int b2_engine_d2h_group(b2_engine_t *e, int group)
{
    if (group < 0 || group >= e->cfg.slots) return -1;
    e->slots[group].host_set = e->slots[group].res_set;
    e->slots[group].polled[e->slots[group].host_set] = 0;
    return 0;
}
int b2_engine_group_result_set(const b2_engine_t *e, int group) { return e->slots[group].host_set; }
int b2_engine_group_done(b2_engine_t *e, int group, int set)
{
    if (e->ncanned) return 1;
    return e->slots[group].polled[set]++ > 0;             /* "not yet" the first time: exercises the wait path */
}
int b2_engine_group_wait(b2_engine_t *e, int group, int set)
{
    struct timespec ts = {0, 200000};
    nanosleep(&ts, NULL);
    e->slots[group].polled[set] = 1;
    return 0;
}
const b2_mbinfo_t *b2_engine_info_set(b2_engine_t *e, int set, int slot) { return e->slots[slot].info[set]; }
const b2_mbinfo_packed_t *b2_engine_info_packed_set(b2_engine_t *e, int set, int slot) { return e->slots[slot].pinfo[set]; }
const uint8_t *b2_engine_packed_set(b2_engine_t *e, int set, int slot, size_t *bytes)
{
    if (bytes) *bytes = e->slots[slot].packed_bytes[set];
    return e->slots[slot].packed[set];
}
