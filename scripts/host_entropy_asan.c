/* Sanitizer driver for the host slice writers (video-encoder_b200/host/b2h_cavlc.c, b2h_cabac.c): reads the cases written by
 * scripts/host_entropy_asan.py (per frame: b2_mbinfo_t[], b2_mbcoef_t[], packed level stream, each in an exact-size heap block),
 * writes every frame with both entropy coders from the dense and from the packed levels (must be identical), and checks that
 * a too small output buffer is reported.  Built with -fsanitize=address,undefined by the script.   usage: driver <case dir> */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b2h_entropy.h"
/* exact-size heap copies so that AddressSanitizer sees any read past the end of info / levels / packed stream */
static void *slurp(const char *p, size_t n) { void *b = malloc(n ? n : 1); FILE *f = fopen(p, "rb"); if (!f || fread(b, 1, n, f) != n) { perror(p); exit(1); } fclose(f); return b; }
int main(int argc, char **argv)
{
    const char *dir = argc > 1 ? argv[1] : ".";
    char mp[512]; snprintf(mp, sizeof mp, "%s/manifest.txt", dir);
    FILE *m = fopen(mp, "r"); if (!m) { perror(mp); return 2; } int idx, w, h, mbw, mbh, qp, ft, t, cabac, deblock, t8; long pn; int n = 0;
    while (fscanf(m, "%d %d %d %d %d %d %d %d %d %d %d %ld", &idx, &w, &h, &mbw, &mbh, &qp, &ft, &t, &cabac, &deblock, &t8, &pn) == 12) {
        char a[600], b[600], c[600]; snprintf(a, sizeof a, "%s/info%d.bin", dir, idx); snprintf(b, sizeof b, "%s/coef%d.bin", dir, idx); snprintf(c, sizeof c, "%s/packed%d.bin", dir, idx);
        const int nmb = mbw * mbh;
        b2_mbinfo_t *info = slurp(a, (size_t)nmb * sizeof(b2_mbinfo_t));
        b2_mbcoef_t *coef = slurp(b, (size_t)nmb * sizeof(b2_mbcoef_t));
        uint8_t *packed = slurp(c, (size_t)pn);
        for (int variant = 0; variant < 2; variant++) {          /* the stream's own entropy mode, then the other one */
            b2h_seq_t s = {w, h, 30, 1, 1, 1, qp, deblock, variant ? !cabac : cabac, t8};
            b2h_entropy_t *e = b2h_entropy_create(mbw, mbh);
            size_t cap = (size_t)nmb * 3072 + 65536; uint8_t *o1 = malloc(cap), *o2 = malloc(cap);
            size_t n1 = b2h_write_slice(e, &s, ft, t, 0, info, coef, o1, cap);
            size_t n2 = b2h_write_slice_packed(e, &s, ft, t, 0, info, packed, (size_t)pn, o2, cap);
            if (!n1 || n1 != n2 || memcmp(o1, o2, n1)) { printf("MISMATCH dense/packed idx %d variant %d (%zu vs %zu)\n", idx, variant, n1, n2); return 1; }
            /* a too small output buffer must be reported, not overrun */
            uint8_t *o3 = malloc(n1 > 8 ? n1 - 5 : 1);
            size_t n3 = b2h_write_slice(e, &s, ft, t, 0, info, coef, o3, n1 > 8 ? n1 - 5 : 1);
            if (n3 != 0) { printf("overflow not reported idx %d\n", idx); return 1; }
            free(o1); free(o2); free(o3); b2h_entropy_destroy(e);
        }
        uint8_t sp[256]; if (!b2h_write_sps(&(b2h_seq_t){w, h, 30, 1, 1, 1, qp, deblock, cabac, t8}, sp, sizeof sp)) return 1;
        free(info); free(coef); free(packed); n++;
    }
    printf("%d frames: dense == packed, overflow reported, no sanitizer finding\n", n);
    return 0;
}
