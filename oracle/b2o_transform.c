/* b2o_transform.c -- ORACLE (test infrastructure only; see b2o.h).
 * 4x4 integer transform, dead-zone quantisation, (normative) dequantisation and inverse
 * transform, the luma-DC 4x4 and chroma-DC 2x2 Hadamards.  In the reference: inside
 * x264_encoder_encode (av_encode.c:970).  Normative half = ITU-T H.264 8.5.9-8.5.12, pinned by
 * the libavcodec decoder drift test; encoder half (forward DCT, quant rounding) = SURVEY.md
 * Appendix A, JM-style quantiser, frozen here. */
#include <stdlib.h>
#include "b2o.h"

const uint8_t b2o_zigzag4x4[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
const uint8_t b2o_blk_x[16] = {0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3};
const uint8_t b2o_blk_y[16] = {0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3};

static const uint16_t quant_mf[6][3] = {
    {13107, 8066, 5243}, {11916, 7490, 4660}, {10082, 6554, 4194},
    {9362, 5825, 3647},  {8192, 5243, 3355},  {7282, 4559, 2893}};
static const uint8_t dequant_v[6][3] = {
    {10, 13, 16}, {11, 14, 18}, {13, 16, 20}, {14, 18, 23}, {16, 20, 25}, {18, 23, 29}};

/* class 0: both indices even, 2: both odd, 1: otherwise */
static inline int pos_class(int i)
{
    int x = i & 3, y = i >> 2;
    return ((x & 1) && (y & 1)) ? 2 : (((x | y) & 1) ? 1 : 0);
}

int b2o_chroma_qp(int qp)
{
    static const uint8_t tab[22] = {29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39};
    return qp < 30 ? qp : tab[qp - 30];
}

void b2o_dct4x4(const int16_t d[16], int16_t w[16])
{
    int t[16];
    for (int y = 0; y < 4; y++) {                       /* rows */
        int s03 = d[y * 4 + 0] + d[y * 4 + 3], d03 = d[y * 4 + 0] - d[y * 4 + 3];
        int s12 = d[y * 4 + 1] + d[y * 4 + 2], d12 = d[y * 4 + 1] - d[y * 4 + 2];
        t[y * 4 + 0] = s03 + s12; t[y * 4 + 1] = 2 * d03 + d12;
        t[y * 4 + 2] = s03 - s12; t[y * 4 + 3] = d03 - 2 * d12;
    }
    for (int x = 0; x < 4; x++) {                       /* columns */
        int s03 = t[x] + t[12 + x], d03 = t[x] - t[12 + x];
        int s12 = t[4 + x] + t[8 + x], d12 = t[4 + x] - t[8 + x];
        w[x] = (int16_t)(s03 + s12); w[4 + x] = (int16_t)(2 * d03 + d12);
        w[8 + x] = (int16_t)(s03 - s12); w[12 + x] = (int16_t)(d03 - 2 * d12);
    }
}

static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* normative 8.5.12.2: rows first, then columns, (x+32)>>6, add to prediction in dst */
static void idct4x4_add32(const int32_t w[16], uint8_t *dst, int pitch)
{
    int t[16];
    for (int y = 0; y < 4; y++) {
        int e0 = w[y * 4 + 0] + w[y * 4 + 2], e1 = w[y * 4 + 0] - w[y * 4 + 2];
        int e2 = (w[y * 4 + 1] >> 1) - w[y * 4 + 3], e3 = w[y * 4 + 1] + (w[y * 4 + 3] >> 1);
        t[y * 4 + 0] = e0 + e3; t[y * 4 + 1] = e1 + e2; t[y * 4 + 2] = e1 - e2; t[y * 4 + 3] = e0 - e3;
    }
    for (int x = 0; x < 4; x++) {
        int e0 = t[x] + t[8 + x], e1 = t[x] - t[8 + x];
        int e2 = (t[4 + x] >> 1) - t[12 + x], e3 = t[4 + x] + (t[12 + x] >> 1);
        int r0 = e0 + e3, r1 = e1 + e2, r2 = e1 - e2, r3 = e0 - e3;
        dst[0 * pitch + x] = (uint8_t)clip255(dst[0 * pitch + x] + ((r0 + 32) >> 6));
        dst[1 * pitch + x] = (uint8_t)clip255(dst[1 * pitch + x] + ((r1 + 32) >> 6));
        dst[2 * pitch + x] = (uint8_t)clip255(dst[2 * pitch + x] + ((r2 + 32) >> 6));
        dst[3 * pitch + x] = (uint8_t)clip255(dst[3 * pitch + x] + ((r3 + 32) >> 6));
    }
}

void b2o_idct4x4_add(const int16_t w[16], uint8_t *dst, int pitch)
{
    int32_t w32[16];
    for (int i = 0; i < 16; i++) w32[i] = w[i];
    idct4x4_add32(w32, dst, pitch);
}

static void hadamard4(const int32_t in[16], int32_t out[16])
{
    int t[16];
    for (int y = 0; y < 4; y++) {
        int s01 = in[y * 4 + 0] + in[y * 4 + 1], d01 = in[y * 4 + 0] - in[y * 4 + 1];
        int s23 = in[y * 4 + 2] + in[y * 4 + 3], d23 = in[y * 4 + 2] - in[y * 4 + 3];
        /* rows of H4: ++++ / ++-- / +--+ / +-+- */
        t[y * 4 + 0] = s01 + s23; t[y * 4 + 1] = s01 - s23; t[y * 4 + 2] = d01 - d23; t[y * 4 + 3] = d01 + d23;
    }
    for (int x = 0; x < 4; x++) {
        int s01 = t[x] + t[4 + x], d01 = t[x] - t[4 + x];
        int s23 = t[8 + x] + t[12 + x], d23 = t[8 + x] - t[12 + x];
        out[x] = s01 + s23; out[4 + x] = s01 - s23; out[8 + x] = d01 - d23; out[12 + x] = d01 + d23;
    }
}

void b2o_hadamard4x4_fwd(const int16_t in[16], int32_t out[16])
{
    int32_t t[16];
    for (int i = 0; i < 16; i++) t[i] = in[i];
    hadamard4(t, out);
    for (int i = 0; i < 16; i++) out[i] = (out[i] + 1) >> 1;
}

void b2o_hadamard4x4_inv(const int32_t in[16], int32_t out[16]) { hadamard4(in, out); }

static inline int deadzone(int qbits, int intra) { return ((1 << qbits) * (intra ? 21 : 11)) >> 6; }

int b2o_quant4x4(const int16_t w[16], int qp, int intra, int skip_dc, int16_t z[16])
{
    int qbits = 15 + qp / 6, f = deadzone(qbits, intra), nnz = 0;
    const uint16_t *mf = quant_mf[qp % 6];
    for (int i = 0; i < 16; i++) {
        if (i == 0 && skip_dc) { z[0] = 0; continue; }
        int a = abs(w[i]);
        int q = (int)(((int64_t)a * mf[pos_class(i)] + f) >> qbits);
        z[i] = (int16_t)(w[i] < 0 ? -q : q);
        nnz += q != 0;
    }
    return nnz;
}

/* DC quantisation shared by luma-DC (16 values) and chroma-DC (4 values) */
static int quant_dc(const int32_t *x, int n, int qp, int intra, int16_t *z)
{
    int qbits = 15 + qp / 6, f = deadzone(qbits, intra), nnz = 0;
    int mf = quant_mf[qp % 6][0];
    for (int i = 0; i < n; i++) {
        int a = abs(x[i]);
        int q = (int)(((int64_t)a * mf + 2 * f) >> (qbits + 1));
        z[i] = (int16_t)(x[i] < 0 ? -q : q);
        nnz += q != 0;
    }
    return nnz;
}

static void dequant4x4_32(const int16_t z[16], int qp, int skip_dc, int32_t w[16])
{
    const uint8_t *v = dequant_v[qp % 6];
    int s = qp / 6;
    for (int i = 0; i < 16; i++) {
        if (i == 0 && skip_dc) continue;               /* caller supplies w[0] */
        int ls = 16 * v[pos_class(i)];
        if (s >= 4) w[i] = (z[i] * ls) * (1 << (s - 4));        /* the standard's "<<" on a possibly negative value */
        else        w[i] = (z[i] * ls + (1 << (3 - s))) >> (4 - s);
    }
}

void b2o_dequant4x4(const int16_t z[16], int qp, int skip_dc, int16_t w[16])
{
    int32_t w32[16];
    w32[0] = w[0];
    dequant4x4_32(z, qp, skip_dc, w32);
    for (int i = 0; i < 16; i++) w[i] = (int16_t)w32[i];
}

/* ---- block-level helpers used by the frame encoder (b2o_encode.c) ---------------------- */

/* residual (src - pred) of one 4x4 block */
static void resid4x4(const uint8_t *src, int sp, const uint8_t *pred, int pp, int16_t d[16])
{
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) d[y * 4 + x] = (int16_t)(src[y * sp + x] - pred[y * pp + x]);
}

/* Plain luma 4x4 (inter or I4x4): recon must hold the prediction on entry.  Levels go to
 * out[16] in zig-zag order.  Returns nnz. */
int b2o_code_luma4x4(const uint8_t *src, int sp, uint8_t *recon, int rp, int qp, int intra, int16_t out[16])
{
    int16_t d[16], w[16], z[16];
    int32_t dq[16];
    resid4x4(src, sp, recon, rp, d);
    b2o_dct4x4(d, w);
    int nnz = b2o_quant4x4(w, qp, intra, 0, z);
    for (int i = 0; i < 16; i++) out[i] = z[b2o_zigzag4x4[i]];
    if (nnz) {
        dequant4x4_32(z, qp, 0, dq);
        idct4x4_add32(dq, recon, rp);
    }
    return nnz;
}

/* I16x16 luma: recon holds the 16x16 prediction.  Fills blk[0..15] (AC, zig-zag, [0]=0) and
 * blk[24] (DC, zig-zag).  Returns nnz mask bits (0-15 AC, 24 DC). */
uint32_t b2o_code_luma16x16(const uint8_t *src, int sp, uint8_t *recon, int rp, int qp, b2_mbcoef_t *c)
{
    int16_t w[16][16], z[16][16], dcin[16], zdc[16];
    int32_t dc[16], dcq[16], dcdq[16];
    uint32_t mask = 0;
    for (int b = 0; b < 16; b++) {
        int16_t d[16];
        int bx = b2o_blk_x[b] * 4, by = b2o_blk_y[b] * 4;
        resid4x4(src + by * sp + bx, sp, recon + by * rp + bx, rp, d);
        b2o_dct4x4(d, w[b]);
        dcin[b2o_blk_y[b] * 4 + b2o_blk_x[b]] = w[b][0];
        if (b2o_quant4x4(w[b], qp, 1, 1, z[b])) mask |= 1u << b;
        for (int i = 0; i < 16; i++) c->blk[b][i] = z[b][b2o_zigzag4x4[i]];
    }
    b2o_hadamard4x4_fwd(dcin, dc);
    if (quant_dc(dc, 16, qp, 1, zdc)) mask |= 1u << 24;
    for (int i = 0; i < 16; i++) c->blk[24][i] = zdc[b2o_zigzag4x4[i]];
    /* decoder side: inverse Hadamard then scale (8.5.10) */
    for (int i = 0; i < 16; i++) dcq[i] = zdc[i];
    b2o_hadamard4x4_inv(dcq, dcdq);
    int ls = 16 * dequant_v[qp % 6][0], s = qp / 6;
    for (int i = 0; i < 16; i++)
        dcdq[i] = s >= 6 ? (dcdq[i] * ls) * (1 << (s - 6)) : (dcdq[i] * ls + (1 << (5 - s))) >> (6 - s);
    int luma_ac = (mask & 0xffffu) != 0;
    for (int b = 0; b < 16; b++) {
        int32_t dq[16];
        int bx = b2o_blk_x[b] * 4, by = b2o_blk_y[b] * 4;
        for (int i = 0; i < 16; i++) dq[i] = 0;
        if (luma_ac) dequant4x4_32(z[b], qp, 1, dq);   /* cbp luma is all-or-nothing for I16x16 */
        dq[0] = dcdq[b2o_blk_y[b] * 4 + b2o_blk_x[b]];
        idct4x4_add32(dq, recon + by * rp + bx, rp);
    }
    return mask;
}

/* One chroma plane of one MB (8x8): recon holds the prediction.  plane: 0 = U, 1 = V.
 * Fills blk[16+4*plane .. +3] (AC) and blk[25][4*plane..] (DC).  Returns nnz mask bits. */
uint32_t b2o_code_chroma8x8(const uint8_t *src, int sp, uint8_t *recon, int rp, int qpc, int intra, int plane,
                            b2_mbcoef_t *c)
{
    int16_t w[4][16], z[4][16], zdc[4];
    int32_t dc[4], f[4];
    uint32_t mask = 0;
    for (int b = 0; b < 4; b++) {
        int16_t d[16];
        int bx = (b & 1) * 4, by = (b >> 1) * 4;
        resid4x4(src + by * sp + bx, sp, recon + by * rp + bx, rp, d);
        b2o_dct4x4(d, w[b]);
        if (b2o_quant4x4(w[b], qpc, intra, 1, z[b])) mask |= 1u << (16 + 4 * plane + b);
        for (int i = 0; i < 16; i++) c->blk[16 + 4 * plane + b][i] = z[b][b2o_zigzag4x4[i]];
    }
    dc[0] = w[0][0] + w[1][0] + w[2][0] + w[3][0];
    dc[1] = w[0][0] - w[1][0] + w[2][0] - w[3][0];
    dc[2] = w[0][0] + w[1][0] - w[2][0] - w[3][0];
    dc[3] = w[0][0] - w[1][0] - w[2][0] + w[3][0];
    if (quant_dc(dc, 4, qpc, intra, zdc)) mask |= 1u << (25 + plane);
    for (int i = 0; i < 4; i++) c->blk[25][4 * plane + i] = zdc[i];
    /* decoder side 8.5.11: inverse 2x2 then ((f*LS) << (qpc/6)) >> 5 */
    f[0] = zdc[0] + zdc[1] + zdc[2] + zdc[3];
    f[1] = zdc[0] - zdc[1] + zdc[2] - zdc[3];
    f[2] = zdc[0] + zdc[1] - zdc[2] - zdc[3];
    f[3] = zdc[0] - zdc[1] - zdc[2] + zdc[3];
    int ls = 16 * dequant_v[qpc % 6][0];
    for (int i = 0; i < 4; i++) f[i] = ((f[i] * ls) * (1 << (qpc / 6))) >> 5;
    int have_ac = (mask >> (16 + 4 * plane) & 15u) != 0;
    for (int b = 0; b < 4; b++) {
        int32_t dq[16];
        int bx = (b & 1) * 4, by = (b >> 1) * 4;
        for (int i = 0; i < 16; i++) dq[i] = 0;
        if (have_ac) dequant4x4_32(z[b], qpc, 1, dq);
        dq[0] = f[b];
        idct4x4_add32(dq, recon + by * rp + bx, rp);
    }
    return mask;
}
