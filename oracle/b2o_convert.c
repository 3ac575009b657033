/* b2o_convert.c -- ORACLE (test infrastructure only; see b2o.h).
 * Closed forms of sws_scale(src fmt -> YUV420P, same size, SWS_FAST_BILINEAR) as the reference
 * calls it at av_encode.c:427-430 / :545-547.  libswscale's source is not under /root/reference;
 * these forms were fitted to, and are pinned bit-exactly by, the live libswscale 9.1.100 bundled
 * with the OpenCV wheel (tests/golden/make_sws_golden.py -> tests/golden/sws_golden.npz):
 *   yuv420p : three strided plane copies
 *   nv12    : Y copy, UV de-interleave
 *   yuyv422 / uyvy422 : Y = every second byte; chroma = vertical average of the two source lines,
 *             rounding (a+b+1)>>1 in the SIMD body (chroma columns < cw & ~7) and truncating
 *             (a+b)>>1 in the scalar tail -- a quirk of that build which the pin preserves. */
#include <string.h>
#include "b2o.h"

int b2o_convert_to_i420(int fmt, int w, int h, const uint8_t *const src[4], const int ss[4],
                        uint8_t *const dst[3], const int ds[3])
{
    int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    if (fmt == B2O_FMT_YUV420P) {
        for (int y = 0; y < h; y++) memcpy(dst[0] + (size_t)y * ds[0], src[0] + (size_t)y * ss[0], w);
        for (int y = 0; y < ch; y++) {
            memcpy(dst[1] + (size_t)y * ds[1], src[1] + (size_t)y * ss[1], cw);
            memcpy(dst[2] + (size_t)y * ds[2], src[2] + (size_t)y * ss[2], cw);
        }
        return 0;
    }
    if (fmt == B2O_FMT_NV12) {
        for (int y = 0; y < h; y++) memcpy(dst[0] + (size_t)y * ds[0], src[0] + (size_t)y * ss[0], w);
        for (int y = 0; y < ch; y++)
            for (int x = 0; x < cw; x++) {
                dst[1][(size_t)y * ds[1] + x] = src[1][(size_t)y * ss[1] + 2 * x];
                dst[2][(size_t)y * ds[2] + x] = src[1][(size_t)y * ss[1] + 2 * x + 1];
            }
        return 0;
    }
    if (fmt == B2O_FMT_YUYV422 || fmt == B2O_FMT_UYVY422) {
        if ((w | h) & 1) return -1;
        int yo = fmt == B2O_FMT_YUYV422 ? 0 : 1, uo = fmt == B2O_FMT_YUYV422 ? 1 : 0, vo = uo + 2;
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
    return -1;
}
