/*
 * b2_mp4.h -- minimal ISO-BMFF writer for one AVC video track, used by tools/b2_encode when the output name ends in ".mp4".
 * It plays the role libmp4v2 has in the reference (MP4Create / MP4AddH264VideoTrack / MP4AddH264Sequence/PictureParameterSet /
 * MP4WriteSample / MP4Close, av_encode.c:600-650, :683-744, :1110-1116) for the video track only -- audio, the
 * reference's other track, is outside this repository's scope.  Samples are the encoder's b_annexb = 0 payloads
 * (4-byte length prefixed NALs); the avcC box comes from b2_avcc_write (include/b2enc.h).
 * Layout: ftyp | mdat (samples; 64-bit `largesize` header patched at close, so payloads beyond 4 GiB stay valid) | moov (sample
 * tables kept in memory until close).  Every write, seek and allocation is checked: a full disk or an out-of-memory condition
 * makes b2_mp4_write_frame / b2_mp4_close return -1 (tools/b2_encode exits non-zero) instead of leaving a corrupt file behind.
 */
#ifndef B2_MP4_H
#define B2_MP4_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b2enc.h"

typedef struct {
    FILE *f;
    int width, height, timescale, delta;
    long mdat_pos;
    uint64_t mdat_bytes;
    uint32_t *sizes; uint8_t *sync; size_t n, cap;
    uint8_t sps[256], pps[256]; int sps_size, pps_size;
    int err;                                 /* sticky: a write, seek or allocation failed */
} b2_mp4_t;

static void mp4_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
typedef struct { uint8_t *d; size_t n, cap; int err; } mp4_buf_t;
static void mb_put(mp4_buf_t *b, const void *p, size_t n)
{
    if (b->err) return;
    if (b->n + n > b->cap) {
        const size_t cap = (b->n + n) * 2 + 256;
        uint8_t *d = (uint8_t *)realloc(b->d, cap);
        if (!d) { b->err = 1; return; }
        b->d = d; b->cap = cap;
    }
    memcpy(b->d + b->n, p, n); b->n += n;
}
static void mb_u32(mp4_buf_t *b, uint32_t v) { uint8_t t[4]; mp4_be32(t, v); mb_put(b, t, 4); }
static void mb_u16(mp4_buf_t *b, unsigned v) { uint8_t t[2] = {(uint8_t)(v >> 8), (uint8_t)v}; mb_put(b, t, 2); }
static void mb_zero(mp4_buf_t *b, size_t n) { static const uint8_t z[64] = {0}; while (n) { size_t k = n > 64 ? 64 : n; mb_put(b, z, k); n -= k; } }
static size_t mb_box(mp4_buf_t *b, const char *type) { size_t at = b->n; mb_u32(b, 0); mb_put(b, type, 4); return at; }
static size_t mb_full(mp4_buf_t *b, const char *type, uint32_t vf) { size_t at = mb_box(b, type); mb_u32(b, vf); return at; }
static void mb_end(mp4_buf_t *b, size_t at) { if (!b->err) mp4_be32(b->d + at, (uint32_t)(b->n - at)); }

static int b2_mp4_open(b2_mp4_t *m, const char *path, int width, int height, int fps_num, int fps_den)
{
    memset(m, 0, sizeof(*m));
    m->f = fopen(path, "wb");
    if (!m->f) return -1;
    m->width = width; m->height = height; m->timescale = fps_num; m->delta = fps_den;
    static const uint8_t ftyp[24] = {0, 0, 0, 24, 'f', 't', 'y', 'p', 'i', 's', 'o', 'm', 0, 0, 2, 0, 'i', 's', 'o', 'm', 'a', 'v', 'c', '1'};
    /* size = 1: the real size follows the type as a 64-bit `largesize` (ISO/IEC 14496-12 4.2), written at close */
    static const uint8_t mdat[16] = {0, 0, 0, 1, 'm', 'd', 'a', 't', 0, 0, 0, 0, 0, 0, 0, 0};
    if (fwrite(ftyp, 1, sizeof(ftyp), m->f) != sizeof(ftyp)) m->err = 1;
    m->mdat_pos = ftell(m->f);
    if (m->mdat_pos < 0 || fwrite(mdat, 1, sizeof(mdat), m->f) != sizeof(mdat)) m->err = 1;
    if (m->err) { fclose(m->f); m->f = NULL; return -1; }
    return 0;
}

/* one encoder output (all NALs of a frame, contiguous, length prefixed), as enc_mp4_write_video_sample walks it (av_encode.c:683-744) */
static int b2_mp4_write_frame(b2_mp4_t *m, const b2_nal_t *nals, int nal_count, int payload_size, int keyframe)
{
    for (int i = 0; i < nal_count; i++) {
        const b2_nal_t *n = &nals[i];
        if (n->i_type == B2_NAL_SPS) { m->sps_size = n->i_payload - 4; memcpy(m->sps, n->p_payload + 4, (size_t)m->sps_size); continue; }
        if (n->i_type == B2_NAL_PPS) { m->pps_size = n->i_payload - 4; memcpy(m->pps, n->p_payload + 4, (size_t)m->pps_size); continue; }
        if (n->i_type == B2_NAL_FILLER) continue;
        const uint8_t *start = n->p_payload;                                   /* everything else: one sample (:733-744) */
        const size_t size = (size_t)payload_size - (size_t)(start - nals[0].p_payload);
        if (m->err || !m->f) return -1;
        if (m->n == m->cap) {
            const size_t cap = m->cap ? 2 * m->cap : 1024;
            uint32_t *sizes = (uint32_t *)realloc(m->sizes, cap * 4);
            if (sizes) m->sizes = sizes;
            uint8_t *sync = (uint8_t *)realloc(m->sync, cap);
            if (sync) m->sync = sync;
            if (!sizes || !sync) { m->err = 1; return -1; }
            m->cap = cap;
        }
        if (fwrite(start, 1, size, m->f) != size) { m->err = 1; return -1; }
        m->sizes[m->n] = (uint32_t)size; m->sync[m->n] = (uint8_t)(keyframe != 0); m->n++;
        m->mdat_bytes += size;
        break;
    }
    return m->err ? -1 : 0;
}

static int b2_mp4_close(b2_mp4_t *m)
{
    if (!m->f) return -1;
    uint8_t avcc[600];
    const int avcc_size = b2_avcc_write(m->sps, m->sps_size, m->pps, m->pps_size, avcc, (int)sizeof(avcc));
    if (avcc_size < 0 || m->err) {
        fprintf(stderr, m->err ? "b2_mp4: write error, the file is incomplete\n" : "b2_mp4: no SPS/PPS seen\n");
        fclose(m->f); m->f = NULL; free(m->sizes); free(m->sync);
        return -1;
    }
    const uint32_t n = (uint32_t)m->n, dur = n * (uint32_t)m->delta;
    static const uint32_t matrix[9] = {0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000};
    mp4_buf_t b = {0};
    size_t moov = mb_box(&b, "moov");
    size_t x = mb_full(&b, "mvhd", 0);
    mb_u32(&b, 0); mb_u32(&b, 0); mb_u32(&b, (uint32_t)m->timescale); mb_u32(&b, dur); mb_u32(&b, 0x10000); mb_u16(&b, 0x100); mb_zero(&b, 10);
    for (int i = 0; i < 9; i++) mb_u32(&b, matrix[i]);
    mb_zero(&b, 24); mb_u32(&b, 2); mb_end(&b, x);
    size_t trak = mb_box(&b, "trak");
    x = mb_full(&b, "tkhd", 3);
    mb_u32(&b, 0); mb_u32(&b, 0); mb_u32(&b, 1); mb_u32(&b, 0); mb_u32(&b, dur); mb_zero(&b, 8); mb_zero(&b, 8);
    for (int i = 0; i < 9; i++) mb_u32(&b, matrix[i]);
    mb_u32(&b, (uint32_t)m->width << 16); mb_u32(&b, (uint32_t)m->height << 16); mb_end(&b, x);
    size_t mdia = mb_box(&b, "mdia");
    x = mb_full(&b, "mdhd", 0); mb_u32(&b, 0); mb_u32(&b, 0); mb_u32(&b, (uint32_t)m->timescale); mb_u32(&b, dur); mb_u16(&b, 0x55c4); mb_u16(&b, 0); mb_end(&b, x);
    x = mb_full(&b, "hdlr", 0); mb_u32(&b, 0); mb_put(&b, "vide", 4); mb_zero(&b, 12); mb_put(&b, "VideoHandler", 13); mb_end(&b, x);
    size_t minf = mb_box(&b, "minf");
    x = mb_full(&b, "vmhd", 1); mb_zero(&b, 8); mb_end(&b, x);
    size_t dinf = mb_box(&b, "dinf"); x = mb_full(&b, "dref", 0); mb_u32(&b, 1); { size_t u = mb_full(&b, "url ", 1); mb_end(&b, u); } mb_end(&b, x); mb_end(&b, dinf);
    size_t stbl = mb_box(&b, "stbl");
    x = mb_full(&b, "stsd", 0); mb_u32(&b, 1);
    {
        size_t avc1 = mb_box(&b, "avc1");
        mb_zero(&b, 6); mb_u16(&b, 1); mb_zero(&b, 16); mb_u16(&b, (unsigned)m->width); mb_u16(&b, (unsigned)m->height);
        mb_u32(&b, 0x480000); mb_u32(&b, 0x480000); mb_u32(&b, 0); mb_u16(&b, 1); mb_zero(&b, 32); mb_u16(&b, 24); mb_u16(&b, 0xffff);
        size_t c = mb_box(&b, "avcC"); mb_put(&b, avcc, (size_t)avcc_size); mb_end(&b, c);
        mb_end(&b, avc1);
    }
    mb_end(&b, x);
    x = mb_full(&b, "stts", 0); mb_u32(&b, 1); mb_u32(&b, n); mb_u32(&b, (uint32_t)m->delta); mb_end(&b, x);
    x = mb_full(&b, "stss", 0);
    { uint32_t ns = 0; for (size_t i = 0; i < m->n; i++) ns += m->sync[i]; mb_u32(&b, ns); for (size_t i = 0; i < m->n; i++) if (m->sync[i]) mb_u32(&b, (uint32_t)i + 1); }
    mb_end(&b, x);
    x = mb_full(&b, "stsc", 0); mb_u32(&b, 1); mb_u32(&b, 1); mb_u32(&b, n); mb_u32(&b, 1); mb_end(&b, x);
    x = mb_full(&b, "stsz", 0); mb_u32(&b, 0); mb_u32(&b, n); for (size_t i = 0; i < m->n; i++) mb_u32(&b, m->sizes[i]); mb_end(&b, x);
    x = mb_full(&b, "stco", 0); mb_u32(&b, 1); mb_u32(&b, (uint32_t)(m->mdat_pos + 16)); mb_end(&b, x);
    mb_end(&b, stbl); mb_end(&b, minf); mb_end(&b, mdia); mb_end(&b, trak); mb_end(&b, moov);
    int err = b.err;
    if (!err && fwrite(b.d, 1, b.n, m->f) != b.n) err = 1;
    uint8_t sz[8];
    const uint64_t total = m->mdat_bytes + 16;
    mp4_be32(sz, (uint32_t)(total >> 32)); mp4_be32(sz + 4, (uint32_t)total);
    if (!err && (fseek(m->f, m->mdat_pos + 8, SEEK_SET) != 0 || fwrite(sz, 1, 8, m->f) != 8)) err = 1;
    if (fclose(m->f) != 0) err = 1;
    m->f = NULL;
    free(b.d); free(m->sizes); free(m->sync);
    if (err) fprintf(stderr, "b2_mp4: write error, the file is incomplete\n");
    return err ? -1 : 0;
}
#endif
