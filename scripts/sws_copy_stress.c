/* Stress / ThreadSanitizer driver for the staging-copy helper pool of video-encoder_b200/host/b2h_sws.c (static functions: the source
 * is included).  Random plane shapes and pitches, contents checked after every picture, pauses that let the helpers fall
 * asleep between pictures.   gcc -O1 -g -fsanitize=thread -Iinclude -Ivideo-encoder_b200/host scripts/sws_copy_stress.c -lpthread */
#include "../video-encoder_b200/host/b2h_sws.c"
/* what b2h_sws.c links against in the product */
void *b2_sws_rt_create(int w, int h, int fmt) { (void)w; (void)h; (void)fmt; return NULL; }
void b2_sws_rt_free(void *rt) { (void)rt; }
int b2_sws_rt_scale(void *rt, const uint8_t *const src[], const int srcStride[], uint8_t *const dst[], const int dstStride[]) { (void)rt; (void)src; (void)srcStride; (void)dst; (void)dstStride; return -1; }
b2h_picrec_t *b2h_picture_find(const uint8_t *p) { (void)p; return NULL; }
uint8_t *b2h_picture_stage(b2h_picrec_t *r, size_t n) { (void)r; (void)n; return NULL; }
#include <time.h>
int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 3000, helpers = argc > 2 ? atoi(argv[2]) : 3;
    copy_pool_t *p = copy_pool_create(helpers);
    if (argc > 3) {                                           /* timing: 1080p I420, tight planes, a fresh source picture every time (32 of them) */
        const int w = 1920, h = 1080;
        const size_t fb = (size_t)w * h * 3 / 2;
        uint8_t *in = malloc(fb * 32), *out = malloc(fb * 2);
        memset(in, 1, fb * 32); memset(out, 2, fb * 2);
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int it = 0; it < iters; it++) {
            const uint8_t *f = in + fb * (size_t)(it % 32);
            uint8_t *o = out + fb * (size_t)(it & 1);
            uint8_t *d[3] = {o, o + (size_t)w * h, o + (size_t)w * h * 5 / 4};
            const uint8_t *sp[3] = {f, f + (size_t)w * h, f + (size_t)w * h * 5 / 4};
            const size_t pitch[3] = {w, w / 2, w / 2}; const int rows[3] = {h, h / 2, h / 2};
            copy_planes(p, 3, d, pitch, sp, pitch, pitch, rows);
        }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double dt = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        printf("1080p I420 staging copy, caller + %d helpers: %.1f us per picture (%.1f GB/s)\n", p ? p->n : 0, dt / iters * 1e6, fb * (double)iters / dt / 1e9);
        copy_pool_destroy(p);
        return 0;
    }
    unsigned s = 12345;
    const size_t cap = (size_t)8 << 20;
    uint8_t *src = malloc(cap), *dst = malloc(cap);
    for (size_t i = 0; i < cap; i++) src[i] = (uint8_t)(i * 2654435761u >> 13);
    for (int it = 0; it < iters; it++) {
        s = s * 1664525u + 1013904223u;
        const int np = 1 + (int)((s >> 8) % 3);
        uint8_t *d[3]; const uint8_t *sp[3]; size_t dpitch[3], spitch[3], rb[3]; int rows[3];
        size_t so = 0, doff = 0;
        for (int k = 0; k < np; k++) {
            s = s * 1664525u + 1013904223u;
            rb[k] = 1 + (s >> 4) % 3000; rows[k] = 1 + (int)((s >> 16) % 700);
            spitch[k] = rb[k] + ((s >> 2) & 1 ? 12 : 0); dpitch[k] = rb[k];
            sp[k] = src + so; d[k] = dst + doff;
            so += spitch[k] * (size_t)rows[k]; doff += dpitch[k] * (size_t)rows[k];
        }
        memset(dst, 0, doff);
        copy_planes(p, np, d, dpitch, sp, spitch, rb, rows);
        for (int k = 0; k < np; k++)
            for (int y = 0; y < rows[k]; y++)
                if (memcmp(d[k] + (size_t)y * dpitch[k], sp[k] + (size_t)y * spitch[k], rb[k])) { printf("MISMATCH iteration %d plane %d row %d\n", it, k, y); return 1; }
        if ((s >> 20) % 50 == 0) { struct timespec ts = {0, 2000000}; nanosleep(&ts, NULL); }     /* helpers go to sleep */
    }
    copy_pool_destroy(p);
    printf("%d pictures copied by the caller + %d helpers: contents exact\n", iters, p ? helpers : 0);
    free(src); free(dst);
    return 0;
}
