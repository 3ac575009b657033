/* b2o_frame.c -- ORACLE (test infrastructure only; see b2o.h).
 * Padded frame handling and the synthetic input generator of SURVEY.md 8(d).
 * Follows what x264_encoder_encode (av_encode.c:970) does to its input picture: copy the
 * caller's I420 planes (x264_picture_alloc'd at av_encode.c:415) into an internal frame whose
 * edges are replicated so that unrestricted motion vectors can read outside the picture. */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "b2o.h"

int b2o_frame_alloc(b2o_frame_t *f, int w, int h)
{
    memset(f, 0, sizeof(*f));
    f->w = w; f->h = h;
    f->w16 = (w + 15) & ~15; f->h16 = (h + 15) & ~15;
    f->mbw = f->w16 >> 4; f->mbh = f->h16 >> 4;
    f->pitch  = f->w16 + 2 * B2_PAD;
    f->pitchc = f->w16 / 2 + 2 * B2_PADC;
    size_t ly = (size_t)f->pitch * (f->h16 + 2 * B2_PAD);
    size_t lc = (size_t)f->pitchc * (f->h16 / 2 + 2 * B2_PADC);
    f->buf[0] = (uint8_t *)calloc(ly, 1);
    f->buf[1] = (uint8_t *)calloc(lc, 1);
    f->buf[2] = (uint8_t *)calloc(lc, 1);
    if (!f->buf[0] || !f->buf[1] || !f->buf[2]) { b2o_frame_free(f); return -1; }
    f->y = f->buf[0] + (size_t)B2_PAD * f->pitch + B2_PAD;
    f->u = f->buf[1] + (size_t)B2_PADC * f->pitchc + B2_PADC;
    f->v = f->buf[2] + (size_t)B2_PADC * f->pitchc + B2_PADC;
    return 0;
}

void b2o_frame_free(b2o_frame_t *f)
{
    for (int i = 0; i < 3; i++) { free(f->buf[i]); f->buf[i] = NULL; }
}

static void extend_plane(uint8_t *p, int pitch, int w, int h, int pad)
{
    for (int y = 0; y < h; y++) {
        uint8_t *row = p + (size_t)y * pitch;
        memset(row - pad, row[0], pad);
        memset(row + w, row[w - 1], pad);
    }
    for (int y = 1; y <= pad; y++) {
        memcpy(p - (size_t)y * pitch - pad, p - pad, w + 2 * pad);
        memcpy(p + (size_t)(h - 1 + y) * pitch - pad, p + (size_t)(h - 1) * pitch - pad, w + 2 * pad);
    }
}

void b2o_frame_extend(b2o_frame_t *f)
{
    extend_plane(f->y, f->pitch, f->w16, f->h16, B2_PAD);
    extend_plane(f->u, f->pitchc, f->w16 / 2, f->h16 / 2, B2_PADC);
    extend_plane(f->v, f->pitchc, f->w16 / 2, f->h16 / 2, B2_PADC);
}

static void load_plane(uint8_t *dst, int dpitch, int w, int h, int cw, int ch,
                       const uint8_t *src, int sstride)
{
    for (int y = 0; y < ch; y++) {
        int sy = y < h ? y : h - 1;
        uint8_t *d = dst + (size_t)y * dpitch;
        memcpy(d, src + (size_t)sy * sstride, w);
        if (cw > w) memset(d + w, d[w - 1], cw - w);
    }
}

void b2o_frame_load(b2o_frame_t *f, const uint8_t *const plane[3], const int stride[3])
{
    int cw = (f->w + 1) >> 1, chh = (f->h + 1) >> 1;
    load_plane(f->y, f->pitch, f->w, f->h, f->w16, f->h16, plane[0], stride[0]);
    load_plane(f->u, f->pitchc, cw, chh, f->w16 / 2, f->h16 / 2, plane[1], stride[1]);
    load_plane(f->v, f->pitchc, cw, chh, f->w16 / 2, f->h16 / 2, plane[2], stride[2]);
    b2o_frame_extend(f);
}

/* ---- synthetic frames: moving gradient + panning texture + per-frame sensor noise ---- */
static inline uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

void b2o_synth_frame(int w, int h, int t, int stream, uint8_t *yp, uint8_t *up, uint8_t *vp)
{
    /* stream 0 pans (+3,+2) px/frame; other streams get other small integer pans */
    int panx = 3 + (stream % 5) - ((stream % 5) > 2 ? 5 : 0);   /* 3,4,5,1,2 */
    int pany = 2 + ((stream / 5) % 3) - (((stream / 5) % 3) > 1 ? 3 : 0);   /* 2,3,1 */
    uint32_t sseed = (uint32_t)stream * 0x9e3779b1U;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t u = (uint32_t)(x + panx * t), v = (uint32_t)(y + pany * t);
            int g = (int)((u + v) & 511u);
            int tri = g < 256 ? g : 511 - g;
            int tex = (int)((((u * 73856093u) ^ (v * 19349663u)) >> 7) & 31u) - 16;
            int n = (int)(hash32((uint32_t)x + (uint32_t)y * 65537u + (uint32_t)t * 0x9e3779b1U + sseed) & 7u) - 4;
            yp[(size_t)y * w + x] = (uint8_t)clip255(16 + ((tri * 219) >> 8) + tex + n);
        }
    int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    for (int y = 0; y < ch; y++)
        for (int x = 0; x < cw; x++) {
            uint32_t u = (uint32_t)(2 * x + panx * t), v = (uint32_t)(2 * y + pany * t);
            up[(size_t)y * cw + x] = (uint8_t)(128 + (int)((u >> 2) & 63u) - 32);
            vp[(size_t)y * cw + x] = (uint8_t)(128 + (int)((v >> 2) & 63u) - 32);
        }
}

double b2o_psnr_y(const b2o_frame_t *a, const b2o_frame_t *b)
{
    double sse = 0;
    for (int y = 0; y < a->h; y++)
        for (int x = 0; x < a->w; x++) {
            int d = a->y[(size_t)y * a->pitch + x] - b->y[(size_t)y * b->pitch + x];
            sse += d * d;
        }
    if (sse == 0) return 99.0;
    return 10.0 * log10(255.0 * 255.0 * a->w * a->h / sse);
}
