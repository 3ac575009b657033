/* b2o_filters.c -- ORACLE (test infrastructure only; see b2o.h).
 * Row N4 of SURVEY.md 8f: the two libavfilter filters the reference's author recommends for DV material
 * ("good DV filter pipeline: hqdn3d,yadif", av_encode.c:35), which the reference runs inside the user-supplied filter
 * graph (av_encode.c:451-517, fed at :962 and drained at :525-560).
 *
 * PARITY UNPINNED: libavfilter is not in this image (the OpenCV wheel bundles avcodec/avformat/avutil/swscale only) and
 * /root/reference holds no filter fixtures, so there is nothing to check these against.  They restate the published
 * algorithms of vf_hqdn3d.c (Daniel Moreno's "high quality 3D denoiser": recursive horizontal / vertical / temporal
 * low-pass with a similarity-weighted coefficient table) and vf_yadif.c (Michael Niedermayer's deinterlacer, mode 0:
 * spatial edge-directed interpolation bounded by a temporal difference check) and ARE the definition the CUDA kernels
 * (csrc/k10_filters.cu) are bit-exact against.  Deliberate simplifications are marked "restatement:". */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "b2o.h"

/* ---- hqdn3d -------------------------------------------------------------------------------------------------------- */
#define LUT_BITS 4
#define LUT_HALF (256 << LUT_BITS)

/* coefficient table for one strength: index = (prev - cur) >> (8 - LUT_BITS) + LUT_HALF, samples in 8.8 fixed point */
void b2o_hqdn3d_coefs(double dist25, int32_t ct[2 * LUT_HALF])
{
    const double d = dist25 < 252.0 ? dist25 : 252.0;
    const double gamma = log(0.25) / log(1.0 - d / 255.0 - 0.00001);
    for (int i = -LUT_HALF; i < LUT_HALF; i++) {
        const double f = (double)((i * (1 << (9 - LUT_BITS))) + (1 << (8 - LUT_BITS)) - 1) / 512.0;   /* midpoint of the bin */
        double simil = 1.0 - fabs(f) / 255.0;
        if (simil < 0) simil = 0;
        ct[LUT_HALF + i] = (int32_t)lrint(pow(simil, gamma) * 256.0 * f);
    }
    if (dist25 <= 0) memset(ct, 0, sizeof(int32_t) * 2 * LUT_HALF);       /* strength 0 switches the pass off */
}

static inline int lowpass(int prev, int cur, const int32_t *ct) { return cur + ct[LUT_HALF + ((prev - cur) >> (8 - LUT_BITS))]; }
static inline int load8(int px) { return (px << 8) + 127; }

/* one plane, one frame.  frame_ant: w*h uint16 state (previous filtered frame, 8.8); first != 0 initialises it from src.
 * restatement: every row starts its horizontal recursion from its own first sample (the original special-cases row 0). */
void b2o_hqdn3d_plane(const uint8_t *src, int sstride, uint8_t *dst, int dstride, int w, int h, uint16_t *frame_ant, int first,
                      const int32_t *spatial, const int32_t *temporal)
{
    int *line_ant = (int *)malloc(sizeof(int) * (size_t)w);
    for (int y = 0; y < h; y++) {
        int pixel_ant = load8(src[(size_t)y * sstride]);
        for (int x = 0; x < w; x++) {
            if (x) pixel_ant = lowpass(pixel_ant, load8(src[(size_t)y * sstride + x]), spatial);          /* horizontal */
            const int v = y ? lowpass(line_ant[x], pixel_ant, spatial) : pixel_ant;                        /* vertical   */
            line_ant[x] = v;
            const int prev = first ? load8(src[(size_t)y * sstride + x]) : frame_ant[(size_t)y * w + x];
            int t = lowpass(prev, v, temporal);                                                            /* temporal   */
            t = t < 0 ? 0 : (t > 65535 ? 65535 : t);
            frame_ant[(size_t)y * w + x] = (uint16_t)t;
            const int o = (t + 128) >> 8;
            dst[(size_t)y * dstride + x] = (uint8_t)(o > 255 ? 255 : o);
        }
    }
    free(line_ant);
}

/* ---- yadif, mode 0 (one frame out per frame in), spatial interlacing check on ----------------------------------------- */
static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int at(const uint8_t *p, int stride, int w, int h, int x, int y)
{
    x = x < 0 ? 0 : (x >= w ? w - 1 : x);                   /* restatement: samples beyond the picture are the edge samples */
    y = y < 0 ? 0 : (y >= h ? h - 1 : y);
    return p[(size_t)y * stride + x];
}

/* one plane.  tff: 1 = top field first.  Lines of the first field are copied from cur, the others are interpolated. */
void b2o_yadif_plane(const uint8_t *prev, const uint8_t *cur, const uint8_t *next, int stride, uint8_t *dst, int dstride, int w, int h,
                     int tff)
{
    const int parity = tff ^ 1;                              /* first output field: lines with (y ^ parity) & 1 are missing */
    const uint8_t *prev2 = prev, *next2 = cur;               /* filter_line is called with parity ^ tff = 1 */
    for (int y = 0; y < h; y++) {
        if (!((y ^ parity) & 1)) { memcpy(dst + (size_t)y * dstride, cur + (size_t)y * stride, (size_t)w); continue; }
        const int ym = y ? y - 1 : y + 1, yp = y + 1 < h ? y + 1 : y - 1;          /* mrefs / prefs with the edge rule of vf_yadif */
        const int limited = y == 1 || y + 2 == h || y == 0 || y + 1 == h;           /* no y+-2 lines: skip the b/f check (mode 2) */
        for (int x = 0; x < w; x++) {
#define CUR(xx, yy) at(cur, stride, w, h, xx, yy)
            const int c = CUR(x, ym), e = CUR(x, yp);
            const int d = (at(prev2, stride, w, h, x, y) + at(next2, stride, w, h, x, y)) >> 1;
            const int td0 = iabs(at(prev2, stride, w, h, x, y) - at(next2, stride, w, h, x, y));
            const int td1 = (iabs(at(prev, stride, w, h, x, ym) - c) + iabs(at(prev, stride, w, h, x, yp) - e)) >> 1;
            const int td2 = (iabs(at(next, stride, w, h, x, ym) - c) + iabs(at(next, stride, w, h, x, yp) - e)) >> 1;
            int diff = imax(imax(td0 >> 1, td1), td2);
            int spatial_pred = (c + e) >> 1;
            int spatial_score = iabs(CUR(x - 1, ym) - CUR(x - 1, yp)) + iabs(c - e) + iabs(CUR(x + 1, ym) - CUR(x + 1, yp)) - 1;
            for (int dir = -1; dir <= 1; dir += 2)          /* CHECK(-1) CHECK(-2), then CHECK(1) CHECK(2): |j| = 2 only after |j| = 1 won */
                for (int k = 1; k <= 2; k++) {
                    const int j = dir * k;
                    const int score = iabs(CUR(x - 1 + j, ym) - CUR(x - 1 - j, yp)) + iabs(CUR(x + j, ym) - CUR(x - j, yp)) +
                                      iabs(CUR(x + 1 + j, ym) - CUR(x + 1 - j, yp));
                    if (score >= spatial_score) break;
                    spatial_score = score;
                    spatial_pred = (CUR(x + j, ym) + CUR(x - j, yp)) >> 1;
                }
            if (!limited) {
                const int b = (at(prev2, stride, w, h, x, y - 2) + at(next2, stride, w, h, x, y - 2)) >> 1;
                const int f = (at(prev2, stride, w, h, x, y + 2) + at(next2, stride, w, h, x, y + 2)) >> 1;
                const int mx = imax(imax(d - e, d - c), imin(b - c, f - e));
                const int mn = imin(imin(d - e, d - c), imax(b - c, f - e));
                diff = imax(imax(diff, mn), -mx);
            }
            if (spatial_pred > d + diff) spatial_pred = d + diff;
            else if (spatial_pred < d - diff) spatial_pred = d - diff;
            dst[(size_t)y * dstride + x] = (uint8_t)spatial_pred;
#undef CUR
        }
    }
}
