/* b2o_pixel.c -- ORACLE (test infrastructure only; see b2o.h).
 * Pixel metrics and the normative H.264 inter-prediction sample interpolation.
 * These are the "linked C functions (pixel SAD/SATD ...)" that BASELINE.json's north_star
 * refers to; in the reference they sit inside libx264 behind av_encode.c:970 and are not
 * available here, so SAD/SATD follow SURVEY.md Appendix A and the interpolation follows
 * ITU-T H.264 8.4.2.2 (pinned by the libavcodec decoder drift test). */
#include <stdlib.h>
#include "b2o.h"

static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

int b2o_mvbits(int v)
{
    if (v == 0) return 1;
    unsigned a = (unsigned)(v < 0 ? -v : v);
    int l = 0;
    while (a >>= 1) l++;
    return 2 * l + 3;
}

int b2o_lambda(int qp)
{
    static const uint8_t tab[52] = {
        1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
        2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 6, 6, 7, 8, 9,
        10, 11, 13, 14, 16, 18, 20, 23, 25, 29, 32, 36, 40, 45, 51, 57,
        64, 72, 81, 91 };
    if (qp < 0) qp = 0;
    if (qp > 51) qp = 51;
    return tab[qp];
}

uint32_t b2o_sad16x16(const uint8_t *a, int pa, const uint8_t *b, int pb)
{
    uint32_t s = 0;
    for (int y = 0; y < 16; y++)
        for (int x = 0; x < 16; x++)
            s += (uint32_t)abs(a[y * pa + x] - b[y * pb + x]);
    return s;
}

uint32_t b2o_satd4x4(const uint8_t *a, int pa, const uint8_t *b, int pb)
{
    int d[16], t[16];
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++)
            d[y * 4 + x] = a[y * pa + x] - b[y * pb + x];
    for (int y = 0; y < 4; y++) {               /* rows */
        int s01 = d[y * 4 + 0] + d[y * 4 + 1], d01 = d[y * 4 + 0] - d[y * 4 + 1];
        int s23 = d[y * 4 + 2] + d[y * 4 + 3], d23 = d[y * 4 + 2] - d[y * 4 + 3];
        t[y * 4 + 0] = s01 + s23; t[y * 4 + 1] = s01 - s23;
        t[y * 4 + 2] = d01 - d23; t[y * 4 + 3] = d01 + d23;
    }
    uint32_t s = 0;
    for (int x = 0; x < 4; x++) {               /* columns */
        int s01 = t[0 + x] + t[4 + x], d01 = t[0 + x] - t[4 + x];
        int s23 = t[8 + x] + t[12 + x], d23 = t[8 + x] - t[12 + x];
        s += (uint32_t)(abs(s01 + s23) + abs(s01 - s23) + abs(d01 - d23) + abs(d01 + d23));
    }
    return s >> 1;
}

uint32_t b2o_satd8x8(const uint8_t *a, int pa, const uint8_t *b, int pb)
{
    uint32_t s = 0;
    for (int y = 0; y < 8; y += 4)
        for (int x = 0; x < 8; x += 4)
            s += b2o_satd4x4(a + y * pa + x, pa, b + y * pb + x, pb);
    return s;
}

uint32_t b2o_satd16x16(const uint8_t *a, int pa, const uint8_t *b, int pb)
{
    uint32_t s = 0;
    for (int y = 0; y < 16; y += 4)
        for (int x = 0; x < 16; x += 4)
            s += b2o_satd4x4(a + y * pa + x, pa, b + y * pb + x, pb);
    return s;
}

/* ---- luma interpolation, H.264 8.4.2.2.1 ---------------------------------------------- */
static inline int tap6(int a, int b, int c, int d, int e, int f)
{
    return a - 5 * b + 20 * c + 20 * d - 5 * e + f;
}
/* unrounded horizontal half sample between (x,y) and (x+1,y) */
static inline int hb1(const uint8_t *r, int p, int x, int y)
{
    const uint8_t *q = r + y * p + x;
    return tap6(q[-2], q[-1], q[0], q[1], q[2], q[3]);
}
/* unrounded vertical half sample between (x,y) and (x,y+1) */
static inline int vh1(const uint8_t *r, int p, int x, int y)
{
    const uint8_t *q = r + y * p + x;
    return tap6(q[-2 * p], q[-p], q[0], q[p], q[2 * p], q[3 * p]);
}
static inline int half_b(const uint8_t *r, int p, int x, int y) { return clip255((hb1(r, p, x, y) + 16) >> 5); }
static inline int half_h(const uint8_t *r, int p, int x, int y) { return clip255((vh1(r, p, x, y) + 16) >> 5); }
static inline int half_j(const uint8_t *r, int p, int x, int y)
{
    int j1 = tap6(hb1(r, p, x, y - 2), hb1(r, p, x, y - 1), hb1(r, p, x, y),
                  hb1(r, p, x, y + 1), hb1(r, p, x, y + 2), hb1(r, p, x, y + 3));
    return clip255((j1 + 512) >> 10);
}

void b2o_mc_luma(const uint8_t *ref, int pitch, int x0, int y0, int mvx, int mvy,
                 int w, int h, uint8_t *dst, int dpitch)
{
    int fx = mvx & 3, fy = mvy & 3;
    int ix = x0 + (mvx >> 2), iy = y0 + (mvy >> 2);
    for (int yy = 0; yy < h; yy++)
        for (int xx = 0; xx < w; xx++) {
            int x = ix + xx, y = iy + yy;
            int G = ref[y * pitch + x];
            int v;
#define AVG(a, b) (((a) + (b) + 1) >> 1)
            switch (fy * 4 + fx) {
            case 0:  v = G; break;
            case 1:  v = AVG(G, half_b(ref, pitch, x, y)); break;                                   /* a */
            case 2:  v = half_b(ref, pitch, x, y); break;                                           /* b */
            case 3:  v = AVG(ref[y * pitch + x + 1], half_b(ref, pitch, x, y)); break;              /* c */
            case 4:  v = AVG(G, half_h(ref, pitch, x, y)); break;                                   /* d */
            case 5:  v = AVG(half_b(ref, pitch, x, y), half_h(ref, pitch, x, y)); break;            /* e */
            case 6:  v = AVG(half_b(ref, pitch, x, y), half_j(ref, pitch, x, y)); break;            /* f */
            case 7:  v = AVG(half_b(ref, pitch, x, y), half_h(ref, pitch, x + 1, y)); break;        /* g */
            case 8:  v = half_h(ref, pitch, x, y); break;                                           /* h */
            case 9:  v = AVG(half_h(ref, pitch, x, y), half_j(ref, pitch, x, y)); break;            /* i */
            case 10: v = half_j(ref, pitch, x, y); break;                                           /* j */
            case 11: v = AVG(half_j(ref, pitch, x, y), half_h(ref, pitch, x + 1, y)); break;        /* k */
            case 12: v = AVG(ref[(y + 1) * pitch + x], half_h(ref, pitch, x, y)); break;            /* n */
            case 13: v = AVG(half_h(ref, pitch, x, y), half_b(ref, pitch, x, y + 1)); break;        /* p */
            case 14: v = AVG(half_j(ref, pitch, x, y), half_b(ref, pitch, x, y + 1)); break;        /* q */
            default: v = AVG(half_h(ref, pitch, x + 1, y), half_b(ref, pitch, x, y + 1)); break;    /* r */
            }
#undef AVG
            dst[yy * dpitch + xx] = (uint8_t)v;
        }
}

/* ---- chroma interpolation, H.264 8.4.2.2.2 -------------------------------------------- */
void b2o_mc_chroma(const uint8_t *ref, int pitch, int x0, int y0, int mvx, int mvy,
                   int w, int h, uint8_t *dst, int dpitch)
{
    int fx = mvx & 7, fy = mvy & 7;
    int ix = x0 + (mvx >> 3), iy = y0 + (mvy >> 3);
    for (int yy = 0; yy < h; yy++)
        for (int xx = 0; xx < w; xx++) {
            const uint8_t *q = ref + (iy + yy) * pitch + ix + xx;
            int A = q[0], B = q[1], C = q[pitch], D = q[pitch + 1];
            dst[yy * dpitch + xx] = (uint8_t)(((8 - fx) * (8 - fy) * A + fx * (8 - fy) * B +
                                               (8 - fx) * fy * C + fx * fy * D + 32) >> 6);
        }
}
