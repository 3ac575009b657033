/* b2o_convert.c -- ORACLE (test infrastructure only; see b2o.h).
 * Closed forms of sws_scale(src fmt -> YUV420P, same size, SWS_FAST_BILINEAR) as the reference
 * calls it at av_encode.c:427-430 / :545-547.  libswscale's source is not under /root/reference;
 * these forms were fitted to, and are pinned bit-exactly by, the live libswscale 9.1.100 bundled
 * with the OpenCV wheel (tests/golden/make_sws_golden.py -> tests/golden/sws_golden.npz):
 *   yuv420p : three strided plane copies
 *   nv12    : Y copy, UV de-interleave
 *   yuyv422 / uyvy422 : Y = every second byte; chroma = vertical average of the two source lines,
 *             rounding (a+b+1)>>1 in the SIMD body (chroma columns < cw & ~7) and truncating
 *             (a+b)>>1 in the scalar tail -- a quirk of that build which the pin preserves.
 *   bgr24   : dedicated converter of that build, bit-exact closed form below (golden vectors).
 * Row N4, TOLERANCE-pinned: rgb24, yuv422p, yuv411p.  For these libswscale runs its generic scaler
 * even at equal sizes, and its SWS_FAST_BILINEAR horizontal pass steps the source position by 65536-20 (x86: +20 with
 * 16-bit wrap) per output pixel: the picture is resampled with a phase that drifts by ~0.0003 px per pixel (0.2 px at
 * 720, wrapping at 1920; measured against the live library: luma "copy" differs by up to 13 at 720 px on smooth content).
 * That artefact is not reproduced; the closed forms below are the drift-free conversion and are pinned to +-1 (+-2 for
 * 4:1:1 chroma) against the live library on pictures <= 70 px wide, where the drift is below rounding
 * (tests/golden/sws_tolerance.npz). */
#include <string.h>
#include "b2o.h"

int b2o_convert_to_i420(int fmt, int w, int h, const uint8_t *const src[4], const int ss[4],
                        uint8_t *const dst[3], const int ds[3])
{
    int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    if (fmt == B2_FMT_YUV420P) {
        for (int y = 0; y < h; y++) memcpy(dst[0] + (size_t)y * ds[0], src[0] + (size_t)y * ss[0], w);
        for (int y = 0; y < ch; y++) {
            memcpy(dst[1] + (size_t)y * ds[1], src[1] + (size_t)y * ss[1], cw);
            memcpy(dst[2] + (size_t)y * ds[2], src[2] + (size_t)y * ss[2], cw);
        }
        return 0;
    }
    if (fmt == B2_FMT_NV12) {
        for (int y = 0; y < h; y++) memcpy(dst[0] + (size_t)y * ds[0], src[0] + (size_t)y * ss[0], w);
        for (int y = 0; y < ch; y++)
            for (int x = 0; x < cw; x++) {
                dst[1][(size_t)y * ds[1] + x] = src[1][(size_t)y * ss[1] + 2 * x];
                dst[2][(size_t)y * ds[2] + x] = src[1][(size_t)y * ss[1] + 2 * x + 1];
            }
        return 0;
    }
    if (fmt == B2_FMT_YUYV422 || fmt == B2_FMT_UYVY422) {
        if ((w | h) & 1) return -1;
        int yo = fmt == B2_FMT_YUYV422 ? 0 : 1, uo = fmt == B2_FMT_YUYV422 ? 1 : 0, vo = uo + 2;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) dst[0][(size_t)y * ds[0] + x] = src[0][(size_t)y * ss[0] + 2 * x + yo];
        int body = cw & ~7;
        for (int y = 0; y < ch; y++) {
            const uint8_t *l0 = src[0] + (size_t)(2 * y) * ss[0], *l1 = l0 + ss[0];
            for (int x = 0; x < cw; x++) {
                int r = x < body ? 1 : 0;
                dst[1][(size_t)y * ds[1] + x] = (uint8_t)((l0[4 * x + uo] + l1[4 * x + uo] + r) >> 1);
                dst[2][(size_t)y * ds[2] + x] = (uint8_t)((l0[4 * x + vo] + l1[4 * x + vo] + r) >> 1);
            }
        }
        return 0;
    }
    if (fmt == B2_FMT_BGR24 || fmt == B2_FMT_RGB24) {
        /* bgr24: libswscale's dedicated bgr24 -> yv12 converter, BIT-EXACT (even widths): 15-bit BT.601 limited-range
         * coefficients, truncating; chroma from the truncated 2x2 mean.  rgb24 runs through the generic scaler there
         * (drifting, see header): same coefficients, rounded, chroma from the 2x2 sum -- tolerance-pinned. */
        if (!b2_fmt_size_ok(fmt, w, h)) return -1;
        const int ro = fmt == B2_FMT_BGR24 ? 2 : 0, bo = 2 - ro, exact = fmt == B2_FMT_BGR24;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const uint8_t *p = src[0] + (size_t)y * ss[0] + 3 * x;
                int v = 8414 * p[ro] + 16519 * p[1] + 3208 * p[bo];
                dst[0][(size_t)y * ds[0] + x] = (uint8_t)(exact ? (v >> 15) + 16 : (v + (16 << 15) + (1 << 14)) >> 15);
            }
        for (int y = 0; y < ch; y++)
            for (int x = 0; x < cw; x++) {
                int y1 = 2 * y + 1 < h ? 2 * y + 1 : h - 1;
                const uint8_t *p0 = src[0] + (size_t)(2 * y) * ss[0] + 6 * x, *p1 = src[0] + (size_t)y1 * ss[0] + 6 * x;
                int r = p0[ro] + p0[3 + ro] + p1[ro] + p1[3 + ro], g = p0[1] + p0[4] + p1[1] + p1[4];
                int b = p0[bo] + p0[3 + bo] + p1[bo] + p1[3 + bo];
                if (exact) {
                    r >>= 2; g >>= 2; b >>= 2;
                    dst[1][(size_t)y * ds[1] + x] = (uint8_t)(((-4865 * r - 9528 * g + 14392 * b) >> 15) + 128);
                    dst[2][(size_t)y * ds[2] + x] = (uint8_t)(((14392 * r - 12061 * g - 2332 * b) >> 15) + 128);
                } else {
                    dst[1][(size_t)y * ds[1] + x] = (uint8_t)((-4865 * r - 9528 * g + 14392 * b + (128 << 17) + (1 << 16)) >> 17);
                    dst[2][(size_t)y * ds[2] + x] = (uint8_t)((14392 * r - 12061 * g - 2332 * b + (128 << 17) + (1 << 16)) >> 17);
                }
            }
        return 0;
    }
    if (fmt == B2_FMT_YUV422P || fmt == B2_FMT_YUV411P) {
        /* planar 4:2:2 / 4:1:1 (DV): luma copied; chroma = rounded mean of the two source lines, 4:1:1 then doubled
         * horizontally (even samples copied, odd samples the rounded mean of their neighbours).  Tolerance-pinned. */
        const int scw = fmt == B2_FMT_YUV422P ? cw : (w + 3) / 4;
        if (!b2_fmt_size_ok(fmt, w, h)) return -1;
        for (int y = 0; y < h; y++) memcpy(dst[0] + (size_t)y * ds[0], src[0] + (size_t)y * ss[0], w);
        for (int p = 1; p < 3; p++)
            for (int y = 0; y < ch; y++) {
                int y1 = 2 * y + 1 < h ? 2 * y + 1 : h - 1;
                const uint8_t *l0 = src[p] + (size_t)(2 * y) * ss[p], *l1 = src[p] + (size_t)y1 * ss[p];
                for (int x = 0; x < cw; x++) {
                    int v;
                    if (fmt == B2_FMT_YUV422P) v = (l0[x] + l1[x] + 1) >> 1;
                    else {
                        int i = x >> 1, i1 = i + 1 < scw ? i + 1 : scw - 1;
                        int a = (l0[i] + l1[i] + 1) >> 1, b = (l0[i1] + l1[i1] + 1) >> 1;
                        v = (x & 1) ? (a + b + 1) >> 1 : a;
                    }
                    dst[p][(size_t)y * ds[p] + x] = (uint8_t)v;
                }
            }
        return 0;
    }
    return -1;
}
