/* Time of the host slice writers (video-encoder_b200/host/b2h_cabac.c, b2h_cavlc.c) per slice, CPU only: reads the cases written by
 * scripts/entropy_timing.py (per frame: b2_mbinfo_t[] and the packed level stream of an oracle encode), writes every slice `reps` times from
 * the packed levels -- the product's path -- and prints mean and minimum time plus a checksum of the bytes (identical bytes before / after
 * a change of the writers).   usage: entropy_timing <case dir> [reps] [cabac 0|1] */
#define _POSIX_C_SOURCE 199309L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "b2h_entropy.h"
static void *slurp(const char *p, size_t n) { void *b = malloc(n ? n : 1); FILE *f = fopen(p, "rb"); if (!f || fread(b, 1, n, f) != n) { perror(p); exit(1); } fclose(f); return b; }
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char **argv)
{
    const char *dir = argc > 1 ? argv[1] : "."; int reps = argc > 2 ? atoi(argv[2]) : 50; int cabac_sel = argc > 3 ? atoi(argv[3]) : 1;
    char mp[512]; snprintf(mp, sizeof mp, "%s/manifest.txt", dir);
    FILE *m = fopen(mp, "r"); if (!m) { perror(mp); return 2; } int idx, w, h, mbw, mbh, qp, ft, t, cabac, deblock, t8; long pn;
    while (fscanf(m, "%d %d %d %d %d %d %d %d %d %d %d %ld", &idx, &w, &h, &mbw, &mbh, &qp, &ft, &t, &cabac, &deblock, &t8, &pn) == 12) {
        char a[600], c[600]; snprintf(a, sizeof a, "%s/info%d.bin", dir, idx); snprintf(c, sizeof c, "%s/packed%d.bin", dir, idx);
        const int nmb = mbw * mbh;
        b2_mbinfo_t *info = slurp(a, (size_t)nmb * sizeof(b2_mbinfo_t));
        uint8_t *packed = slurp(c, (size_t)pn);
        b2h_seq_t s = {w, h, 30, 1, 1, 1, qp, deblock, cabac_sel, t8, 0, 0};
        b2h_entropy_t *e = b2h_entropy_create(mbw, mbh);
        size_t cap = (size_t)nmb * 3072 + 65536; uint8_t *o1 = malloc(cap);
        size_t n1 = 0; unsigned long sum = 0;
        double best = 1e9, t0 = now();
        for (int r = 0; r < reps; r++) { double a0 = now(); n1 = b2h_write_slice_packed(e, &s, ft, t, 0, info, packed, (size_t)pn, o1, cap); double a1 = now(); if (a1 - a0 < best) best = a1 - a0; }
        double t1 = now();
        for (size_t i = 0; i < n1; i++) sum = sum * 31 + o1[i];
        printf("frame %d type %d: %zu bytes, %.3f ms per slice (min %.3f), packed %ld, sum %lx\n", idx, ft, n1, (t1 - t0) / reps * 1e3, best * 1e3, pn, sum);
        free(o1); b2h_entropy_destroy(e); free(info); free(packed);
    }
    return 0;
}
