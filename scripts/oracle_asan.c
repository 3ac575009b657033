/* Sanitizer driver for the C oracle (oracle/b2o_*.c): random picture sizes, QPs, search ranges and tool sets, random / flat /
 * panning content, three frames each (I, P, P), everything built with -fsanitize=address,undefined -- a read outside the
 * padded planes or an overflow in the restated arithmetic would make "bit-exact against the oracle" meaningless.
 * Built and run by scripts/oracle_asan.sh.   usage: oracle_asan [cases] [seed] */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b2o.h"

static unsigned rs = 1;
static unsigned rnd(void) { rs = rs * 1664525u + 1013904223u; return rs >> 8; }

int main(int argc, char **argv)
{
    const int cases = argc > 1 ? atoi(argv[1]) : 60;
    rs = argc > 2 ? (unsigned)atoi(argv[2]) : 1;
    unsigned long long sum = 0;
    for (int c = 0; c < cases; c++) {
        const int w = 16 + 2 * (int)(rnd() % 90), h = 16 + 2 * (int)(rnd() % 70);
        b2o_params_t p = {10 + (int)(rnd() % 42), (rnd() & 1) ? 32 : 16, (rnd() % 8) != 0, (rnd() % 8) != 0, (int)(rnd() & 1), (int)(rnd() & 1), (int)(rnd() % 3)};
        if (!p.subpel) p.partitions = 0;
        const int kind = (int)(rnd() % 3);
        b2o_frame_t cur, rec[2];
        if (b2o_frame_alloc(&cur, w, h) || b2o_frame_alloc(&rec[0], w, h) || b2o_frame_alloc(&rec[1], w, h)) return 2;
        const int nmb = cur.mbw * cur.mbh, cw = (w + 1) / 2, ch = (h + 1) / 2;
        b2_mbinfo_t *info = (b2_mbinfo_t *)malloc((size_t)nmb * sizeof(*info));
        b2_mbcoef_t *coef = (b2_mbcoef_t *)malloc((size_t)nmb * sizeof(*coef));
        b2_mv_t *pmv = (b2_mv_t *)malloc((size_t)nmb * sizeof(*pmv));
        uint8_t *y = (uint8_t *)malloc((size_t)w * h), *u = (uint8_t *)malloc((size_t)cw * ch), *v = (uint8_t *)malloc((size_t)cw * ch);
        for (int t = 0; t < 3; t++) {
            if (kind == 0) b2o_synth_frame(w, h, t, c % 7, y, u, v);
            else if (kind == 1) { for (int i = 0; i < w * h; i++) y[i] = (uint8_t)rnd(); for (int i = 0; i < cw * ch; i++) { u[i] = (uint8_t)rnd(); v[i] = (uint8_t)rnd(); } }
            else { memset(y, (int)(rnd() & 255), (size_t)w * h); memset(u, (int)(rnd() & 255), (size_t)cw * ch); memset(v, (int)(rnd() & 255), (size_t)cw * ch); }
            const uint8_t *pl[3] = {y, u, v}; const int st[3] = {w, cw, cw};
            b2o_frame_load(&cur, pl, st);
            b2o_encode_frame(&p, t == 0 ? B2_FRAME_I : B2_FRAME_P, &cur, t == 0 ? NULL : &rec[(t + 1) & 1], &rec[t & 1], t > 1 ? pmv : NULL, info, coef);
            for (int i = 0; i < nmb; i++) { pmv[i].x = info[i].mvx; pmv[i].y = info[i].mvy; sum += info[i].cost + info[i].cbp; }
        }
        free(info); free(coef); free(pmv); free(y); free(u); free(v);
        b2o_frame_free(&cur); b2o_frame_free(&rec[0]); b2o_frame_free(&rec[1]);
    }
    printf("%d cases x 3 frames through the oracle encode stage, checksum %llu, no sanitizer finding\n", cases, sum);
    return 0;
}
