/* b2o_transform8.c -- ORACLE (test infrastructure only; see b2o.h).
 * 8x8 integer transform path of the High profile: forward 8x8 DCT and dead-zone quantiser
 * (encoder side, frozen here; x264's dct8 butterflies and deadzone form), normative 8x8 scaling (8.5.13, flat
 * matrices) and inverse transform (8.5.13), the 8x8 zig-zag scan, and SA8D (8x8 Hadamard SATD) which decides
 * between the 4x4 and the 8x8 transform of an inter macroblock the way x264's non-RD analysis does
 * (x264_mb_analyse_transform: sa8d 16x16 < satd 16x16).  In the reference all of it is inside
 * x264_encoder_encode (av_encode.c:970).  The normative half is pinned by the libavcodec decoder drift test. */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "b2o.h"

uint8_t b2o_zigzag8x8[64];                 /* scan pos -> raster index (y*8+x); built on first use */
static pthread_once_t zz8_once = PTHREAD_ONCE_INIT;   /* bench.py's cpu_baseline runs one GOP per host thread */

static void zz8_init(void)
{
    /* classic zig-zag (Figure 8-8, 8x8 frame scan): anti-diagonals, alternating direction */
    int n = 0;
    for (int d = 0; d < 15; d++) {
        if (d & 1) { for (int y = d < 8 ? 0 : d - 7; y <= (d < 8 ? d : 7); y++) b2o_zigzag8x8[n++] = (uint8_t)(y * 8 + (d - y)); }
        else       { for (int x = d < 8 ? 0 : d - 7; x <= (d < 8 ? d : 7); x++) b2o_zigzag8x8[n++] = (uint8_t)((d - x) * 8 + x); }
    }
}
const uint8_t *b2o_zigzag8(void) { pthread_once(&zz8_once, zz8_init); return b2o_zigzag8x8; }

/* normAdjust8x8 position classes (8.5.9): [y&3][x&3] */
static const uint8_t cls8[16] = {0, 3, 4, 3, 3, 1, 5, 1, 4, 5, 2, 5, 3, 1, 5, 1};
static const uint16_t quant8_mf[6][6] = {
    {13107, 11428, 20972, 12222, 16777, 15481}, {11916, 10826, 19174, 11058, 14980, 14290},
    {10082, 8943, 15978, 9675, 12710, 11985},   {9362, 8228, 14913, 8931, 11984, 11259},
    {8192, 7346, 13159, 7740, 10486, 9777},     {7282, 6428, 11570, 6830, 9118, 8640}};
static const uint8_t dequant8_v[6][6] = {
    {20, 18, 32, 19, 25, 24}, {22, 19, 35, 21, 28, 26}, {26, 23, 42, 24, 33, 31},
    {28, 25, 45, 26, 35, 33}, {32, 28, 51, 30, 40, 38}, {36, 32, 58, 34, 46, 43}};
static inline int pos_class8(int i) { return cls8[((i >> 3) & 3) * 4 + (i & 3)]; }

static void fdct8_1d(const int s[8], int d[8])
{
    const int s07 = s[0] + s[7], s16 = s[1] + s[6], s25 = s[2] + s[5], s34 = s[3] + s[4];
    const int a0 = s07 + s34, a1 = s16 + s25, a2 = s07 - s34, a3 = s16 - s25;
    const int d07 = s[0] - s[7], d16 = s[1] - s[6], d25 = s[2] - s[5], d34 = s[3] - s[4];
    const int a4 = d16 + d25 + (d07 + (d07 >> 1));
    const int a5 = d07 - d34 - (d25 + (d25 >> 1));
    const int a6 = d07 + d34 - (d16 + (d16 >> 1));
    const int a7 = d16 - d25 + (d34 + (d34 >> 1));
    d[0] = a0 + a1; d[1] = a4 + (a7 >> 2); d[2] = a2 + (a3 >> 1); d[3] = a5 + (a6 >> 2);
    d[4] = a0 - a1; d[5] = a6 - (a5 >> 2); d[6] = (a2 >> 1) - a3; d[7] = (a4 >> 2) - a7;
}

/* forward 8x8: columns first, then rows (x264's order); raster in, raster out (w[v*8+u], u horizontal) */
void b2o_dct8x8(const int16_t d[64], int32_t w[64])
{
    int t[64], s[8], o[8];
    for (int x = 0; x < 8; x++) {
        for (int y = 0; y < 8; y++) s[y] = d[y * 8 + x];
        fdct8_1d(s, o);
        for (int y = 0; y < 8; y++) t[y * 8 + x] = o[y];
    }
    for (int y = 0; y < 8; y++) {
        fdct8_1d(t + y * 8, o);
        for (int x = 0; x < 8; x++) w[y * 8 + x] = o[x];
    }
}

static void idct8_1d(const int d[8], int o[8])
{
    const int a0 = d[0] + d[4], a2 = d[0] - d[4], a4 = (d[2] >> 1) - d[6], a6 = d[2] + (d[6] >> 1);
    const int b0 = a0 + a6, b2 = a2 + a4, b4 = a2 - a4, b6 = a0 - a6;
    const int a1 = -d[3] + d[5] - d[7] - (d[7] >> 1);
    const int a3 = d[1] + d[7] - d[3] - (d[3] >> 1);
    const int a5 = -d[1] + d[7] + d[5] + (d[5] >> 1);
    const int a7 = d[3] + d[5] + d[1] + (d[1] >> 1);
    const int b1 = a1 + (a7 >> 2), b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5, b7 = a7 - (a1 >> 2);
    o[0] = b0 + b7; o[1] = b2 + b5; o[2] = b4 + b3; o[3] = b6 + b1;
    o[4] = b6 - b1; o[5] = b4 - b3; o[6] = b2 - b5; o[7] = b0 - b7;
}

static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* normative 8.5.13: each row, then each column, then (x + 32) >> 6, added to the prediction in dst */
void b2o_idct8x8_add(const int32_t w[64], uint8_t *dst, int pitch)
{
    int t[64], s[8], o[8];
    for (int y = 0; y < 8; y++) {
        for (int x = 0; x < 8; x++) s[x] = w[y * 8 + x];
        idct8_1d(s, o);
        for (int x = 0; x < 8; x++) t[y * 8 + x] = o[x];
    }
    for (int x = 0; x < 8; x++) {
        for (int y = 0; y < 8; y++) s[y] = t[y * 8 + x];
        idct8_1d(s, o);
        for (int y = 0; y < 8; y++) dst[y * pitch + x] = (uint8_t)clip255(dst[y * pitch + x] + ((o[y] + 32) >> 6));
    }
}

int b2o_quant8x8(const int32_t w[64], int qp, int intra, int16_t z[64])
{
    const int qbits = 16 + qp / 6, f = ((1 << qbits) * (intra ? 21 : 11)) >> 6;
    int nnz = 0;
    for (int i = 0; i < 64; i++) {
        const int a = abs(w[i]);
        const int q = (int)(((int64_t)a * quant8_mf[qp % 6][pos_class8(i)] + f) >> qbits);
        z[i] = (int16_t)(w[i] < 0 ? -q : q);
        nnz += q != 0;
    }
    return nnz;
}

void b2o_dequant8x8(const int16_t z[64], int qp, int32_t w[64])
{
    const int s = qp / 6;
    for (int i = 0; i < 64; i++) {
        const int ls = 16 * dequant8_v[qp % 6][pos_class8(i)];
        w[i] = s >= 6 ? (z[i] * ls) * (1 << (s - 6)) : (z[i] * ls + (1 << (5 - s))) >> (6 - s);
    }
}

/* One luma 8x8 block: recon holds the prediction on entry and the reconstruction on return.  out[64] = levels in
 * 8x8 zig-zag order.  Returns a 4-bit mask: bit k set when the k-th interleaved quarter (levels 4i+k, the CAVLC
 * split of 7.3.5.3.2) has a non-zero level. */
int b2o_code_luma8x8(const uint8_t *src, int sp, uint8_t *recon, int rp, int qp, int intra, int16_t out[64])
{
    int16_t d[64], z[64];
    int32_t w[64];
    const uint8_t *zz = b2o_zigzag8();
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) d[y * 8 + x] = (int16_t)(src[y * sp + x] - recon[y * rp + x]);
    b2o_dct8x8(d, w);
    const int nnz = b2o_quant8x8(w, qp, intra, z);
    int mask = 0;
    for (int i = 0; i < 64; i++) {
        out[i] = z[zz[i]];
        if (out[i]) mask |= 1 << (i & 3);
    }
    if (nnz) {
        b2o_dequant8x8(z, qp, w);
        b2o_idct8x8_add(w, recon, rp);
    }
    return mask;
}

/* sum |H8 D H8^T| of one 8x8 difference block (unnormalised) */
static uint32_t hadamard8_abs(const uint8_t *a, int pa, const uint8_t *b, int pb)
{
    int m[64];
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) m[y * 8 + x] = a[y * pa + x] - b[y * pb + x];
    for (int pass = 0; pass < 2; pass++) {
        const int st = pass ? 8 : 1, ln = pass ? 1 : 8;          /* rows, then columns */
        for (int l = 0; l < 8; l++) {
            int *p = m + l * ln;
            for (int h = 1; h < 8; h <<= 1)
                for (int i = 0; i < 8; i += 2 * h)
                    for (int j = i; j < i + h; j++) {
                        const int u = p[j * st], v = p[(j + h) * st];
                        p[j * st] = u + v; p[(j + h) * st] = u - v;
                    }
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < 64; i++) s += (uint32_t)abs(m[i]);
    return s;
}

/* SA8D of one 8x8 block: (sum |H8 D H8^T| + 2) >> 2 (x264's sa8d_8x8 normalisation) */
uint32_t b2o_sa8d8x8(const uint8_t *a, int pa, const uint8_t *b, int pb) { return (hadamard8_abs(a, pa, b, pb) + 2) >> 2; }

/* SA8D of a 16x16 block: (sum over the four 8x8 Hadamards + 2) >> 2 */
uint32_t b2o_sa8d16x16(const uint8_t *a, int pa, const uint8_t *b, int pb)
{
    uint32_t s = 0;
    for (int y = 0; y < 16; y += 8)
        for (int x = 0; x < 16; x += 8) s += hadamard8_abs(a + y * pa + x, pa, b + y * pb + x, pb);
    return (s + 2) >> 2;
}
