/* Host-stage ceiling of the drop-in (CPU only): the reference's per-frame pair b2_sws_scale(picture -> pic_in) + b2_encoder_encode(pic_in)
 * (av_encode.c:545-547, :970) against the zero-latency mock engine (tests/mock/mock_engine.c, B2_MOCK_CANNED), i.e. the product's host code --
 * staging copy, GOP bookkeeping, per-GPU threads, result copies, entropy workers, display-order fifo -- with GPUs that take no time.
 * Built and run by scripts/host_ceiling.py.   usage: host_ceiling WIDTH HEIGHT FRAMES [DEVICES] [SLOTS] */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "b2enc.h"
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: host_ceiling WIDTH HEIGHT FRAMES [DEVICES] [SLOTS]\n"); return 2; }
    const int w = atoi(argv[1]), h = atoi(argv[2]), frames = atoi(argv[3]), devices = argc > 4 ? atoi(argv[4]) : 1, slots = argc > 5 ? atoi(argv[5]) : 0;
    b2_param_t prm;
    if (b2_param_default_preset(&prm, "slow", "film")) return 3;
    prm.i_width = w; prm.i_height = h; prm.i_fps_num = 60; prm.i_fps_den = 1; prm.rc.i_rc_method = B2_RC_CRF; prm.rc.f_rf_constant = 26;
    prm.b_annexb = 1; prm.i_devices = devices;
    if (slots) prm.i_gop_slots = slots;
    b2_t *enc = b2_encoder_open(&prm);
    b2_picture_t pic_in, pic_out;
    if (!enc || b2_picture_alloc(&pic_in, B2_CSP_I420, w, h)) { fprintf(stderr, "open failed\n"); return 4; }
    b2_sws_context_t *sws = b2_sws_getContext(w, h, B2_FMT_YUV420P, w, h, B2_FMT_YUV420P, B2_SWS_FAST_BILINEAR, NULL, NULL, NULL);
    if (!sws) return 5;
    const int cw = (w + 1) / 2, ch = (h + 1) / 2, ring = 32;
    const size_t fb = (size_t)w * h + 2 * (size_t)cw * ch;
    uint8_t *in = malloc(fb * ring);                                   /* a decoder's picture pool: pageable memory, > last-level cache */
    for (size_t i = 0; i < fb * ring; i++) in[i] = (uint8_t)(i * 2654435761u >> 11);
    b2_nal_t *nals; int nn; long out = 0; size_t bytes = 0;
    const double t0 = now();
    for (int t = 0; t < frames; t++) {
        const uint8_t *raw = in + fb * (size_t)(t % ring);
        const uint8_t *src[4] = {raw, raw + (size_t)w * h, raw + (size_t)w * h + (size_t)cw * ch, NULL};
        const int stride[4] = {w, cw, cw, 0};
        if (b2_sws_scale(sws, src, stride, 0, h, pic_in.img.plane, pic_in.img.i_stride) != h) return 6;
        pic_in.i_type = B2_TYPE_AUTO; pic_in.i_pts = t;
        const int n = b2_encoder_encode(enc, &nals, &nn, &pic_in, &pic_out);
        if (n < 0) return 7;
        if (n > 0) { out++; bytes += (size_t)n; }
    }
    const double t1 = now();
    while (b2_encoder_delayed_frames(enc) > 0) {
        const int n = b2_encoder_encode(enc, &nals, &nn, NULL, &pic_out);
        if (n <= 0) break;
        out++; bytes += (size_t)n;
    }
    const double t2 = now();
    printf("%dx%d, %d pretend GPU(s), %d GOP slots each: %d frames in, %ld out, %zu bytes; producer loop %.3f s = %.0f frames/s, with drain %.3f s = %.0f frames/s\n",
           w, h, devices, prm.i_gop_slots, frames, out, bytes, t1 - t0, frames / (t1 - t0), t2 - t0, out / (t2 - t0));
    b2_sws_freeContext(sws); b2_picture_clean(&pic_in); b2_encoder_close(enc); free(in);
    return out == frames ? 0 : 8;
}
