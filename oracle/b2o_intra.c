/* b2o_intra.c -- ORACLE (test infrastructure only; see b2o.h).
 * Intra predictors (normative, ITU-T H.264 8.3.1.2 / 8.3.3 / 8.3.4; pinned by the libavcodec
 * decoder drift test) and the source-pixel intra analysis (encoder choice, frozen here).
 * In the reference this lives inside x264_encoder_encode (av_encode.c:970).
 *
 * Analysis rule ("decide on source pixels, reconstruct in wavefront
 * order"): mode costs are SATD(source block, prediction built from *source* neighbours) +
 * lambda * bits, so every MB and every 4x4 block is decided independently; reconstruction
 * later applies the chosen modes to *reconstructed* neighbours, as the decoder will. */
#include <stdlib.h>
#include <string.h>
#include "b2o.h"

static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

int b2o_i16_mode_ok(int mode, int avail)
{
    switch (mode) {
    case B2_I16_V: return (avail & B2O_AV_T) != 0;
    case B2_I16_H: return (avail & B2O_AV_L) != 0;
    case B2_I16_DC: return 1;
    default: return (avail & (B2O_AV_L | B2O_AV_T | B2O_AV_TL)) == (B2O_AV_L | B2O_AV_T | B2O_AV_TL);
    }
}

int b2o_i4_mode_ok(int mode, int avail)
{
    const int LT = B2O_AV_L | B2O_AV_T | B2O_AV_TL;
    switch (mode) {
    case B2_I4_V: case B2_I4_DDL: case B2_I4_VL: return (avail & B2O_AV_T) != 0;
    case B2_I4_H: case B2_I4_HU: return (avail & B2O_AV_L) != 0;
    case B2_I4_DC: return 1;
    default: return (avail & LT) == LT;      /* DDR, VR, HD */
    }
}

void b2o_pred16x16(int mode, const uint8_t *p, int pitch, int avail, uint8_t dst[256])
{
    const uint8_t *top = p - pitch;
    if (mode == B2_I16_V) {
        for (int y = 0; y < 16; y++) memcpy(dst + y * 16, top, 16);
    } else if (mode == B2_I16_H) {
        for (int y = 0; y < 16; y++) memset(dst + y * 16, p[y * pitch - 1], 16);
    } else if (mode == B2_I16_DC) {
        int s = 0, dc;
        if (avail & B2O_AV_T) for (int i = 0; i < 16; i++) s += top[i];
        if (avail & B2O_AV_L) for (int i = 0; i < 16; i++) s += p[i * pitch - 1];
        if ((avail & B2O_AV_T) && (avail & B2O_AV_L)) dc = (s + 16) >> 5;
        else if (avail & (B2O_AV_T | B2O_AV_L)) dc = (s + 8) >> 4;
        else dc = 128;
        memset(dst, dc, 256);
    } else {
        int H = 0, V = 0;
        for (int i = 0; i < 8; i++) {
            H += (i + 1) * (top[8 + i] - top[6 - i]);                       /* top[-1] = top-left */
            V += (i + 1) * (p[(8 + i) * pitch - 1] - p[(6 - i) * pitch - 1]);
        }
        int a = 16 * (p[15 * pitch - 1] + top[15]);
        int b = (5 * H + 32) >> 6, c = (5 * V + 32) >> 6;
        for (int y = 0; y < 16; y++)
            for (int x = 0; x < 16; x++)
                dst[y * 16 + x] = (uint8_t)clip255((a + b * (x - 7) + c * (y - 7) + 16) >> 5);
    }
}

void b2o_pred8x8c(int mode, const uint8_t *p, int pitch, int avail, uint8_t dst[64])
{
    const uint8_t *top = p - pitch;
    int T = (avail & B2O_AV_T) != 0, L = (avail & B2O_AV_L) != 0;
    if (mode == B2_IC_V) {
        for (int y = 0; y < 8; y++) memcpy(dst + y * 8, top, 8);
    } else if (mode == B2_IC_H) {
        for (int y = 0; y < 8; y++) memset(dst + y * 8, p[y * pitch - 1], 8);
    } else if (mode == B2_IC_DC) {
        int t0 = 0, t1 = 0, l0 = 0, l1 = 0, dc[4];
        if (T) for (int i = 0; i < 4; i++) { t0 += top[i]; t1 += top[4 + i]; }
        if (L) for (int i = 0; i < 4; i++) { l0 += p[i * pitch - 1]; l1 += p[(4 + i) * pitch - 1]; }
        dc[0] = (T && L) ? (t0 + l0 + 4) >> 3 : T ? (t0 + 2) >> 2 : L ? (l0 + 2) >> 2 : 128;
        dc[1] = T ? (t1 + 2) >> 2 : L ? (l0 + 2) >> 2 : 128;
        dc[2] = L ? (l1 + 2) >> 2 : T ? (t0 + 2) >> 2 : 128;
        dc[3] = (T && L) ? (t1 + l1 + 4) >> 3 : T ? (t1 + 2) >> 2 : L ? (l1 + 2) >> 2 : 128;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) dst[y * 8 + x] = (uint8_t)dc[(y >> 2) * 2 + (x >> 2)];
    } else {
        int H = 0, V = 0;
        for (int i = 0; i < 4; i++) {
            H += (i + 1) * (top[4 + i] - top[2 - i]);
            V += (i + 1) * (p[(4 + i) * pitch - 1] - p[(2 - i) * pitch - 1]);
        }
        int a = 16 * (p[7 * pitch - 1] + top[7]);
        int b = (34 * H + 32) >> 6, c = (34 * V + 32) >> 6;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++)
                dst[y * 8 + x] = (uint8_t)clip255((a + b * (x - 3) + c * (y - 3) + 16) >> 5);
    }
}

void b2o_pred4x4(int mode, const uint8_t *p, int pitch, int avail, uint8_t dst[16])
{
    /* E[0] = top-left M, E[1..8] = top T0..T7 ; Lf[0] = M, Lf[1..4] = left L0..L3 */
    int Tb[9], Lb[5];
    int *T = Tb + 1, *L = Lb + 1;
    const uint8_t *top = p - pitch;
    int M = (avail & B2O_AV_TL) ? top[-1] : 128;
    T[-1] = L[-1] = M;
    for (int i = 0; i < 4; i++) {
        T[i] = (avail & B2O_AV_T) ? top[i] : 128;
        L[i] = (avail & B2O_AV_L) ? p[i * pitch - 1] : 128;
    }
    for (int i = 4; i < 8; i++) T[i] = (avail & B2O_AV_TR) ? top[i] : T[3];
#define F3(a, b, c) (((a) + 2 * (b) + (c) + 2) >> 2)
#define F2(a, b) (((a) + (b) + 1) >> 1)
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) {
            int v;
            switch (mode) {
            case B2_I4_V: v = T[x]; break;
            case B2_I4_H: v = L[y]; break;
            case B2_I4_DC: {
                int s = 0;
                int hasT = (avail & B2O_AV_T) != 0, hasL = (avail & B2O_AV_L) != 0;
                if (hasT) s += T[0] + T[1] + T[2] + T[3];
                if (hasL) s += L[0] + L[1] + L[2] + L[3];
                v = (hasT && hasL) ? (s + 4) >> 3 : (hasT || hasL) ? (s + 2) >> 2 : 128;
                break;
            }
            case B2_I4_DDL:
                v = (x == 3 && y == 3) ? (T[6] + 3 * T[7] + 2) >> 2 : F3(T[x + y], T[x + y + 1], T[x + y + 2]);
                break;
            case B2_I4_DDR:
                if (x > y) v = F3(T[x - y - 2], T[x - y - 1], T[x - y]);
                else if (x < y) v = F3(L[y - x - 2], L[y - x - 1], L[y - x]);
                else v = F3(T[0], M, L[0]);
                break;
            case B2_I4_VR: {
                int z = 2 * x - y;
                if (z >= 0 && !(z & 1)) v = F2(T[x - (y >> 1) - 1], T[x - (y >> 1)]);
                else if (z >= 0) v = F3(T[x - (y >> 1) - 2], T[x - (y >> 1) - 1], T[x - (y >> 1)]);
                else if (z == -1) v = F3(L[0], M, T[0]);
                else v = F3(L[y - 1], L[y - 2], L[y - 3]);
                break;
            }
            case B2_I4_HD: {
                int z = 2 * y - x;
                if (z >= 0 && !(z & 1)) v = F2(L[y - (x >> 1) - 1], L[y - (x >> 1)]);
                else if (z >= 0) v = F3(L[y - (x >> 1) - 2], L[y - (x >> 1) - 1], L[y - (x >> 1)]);
                else if (z == -1) v = F3(L[0], M, T[0]);
                else v = F3(T[x - 1], T[x - 2], T[x - 3]);
                break;
            }
            case B2_I4_VL:
                v = (y & 1) ? F3(T[x + (y >> 1)], T[x + (y >> 1) + 1], T[x + (y >> 1) + 2])
                            : F2(T[x + (y >> 1)], T[x + (y >> 1) + 1]);
                break;
            default: {                                   /* HU */
                int z = x + 2 * y;
                if (z > 5) v = L[3];
                else if (z == 5) v = (L[2] + 3 * L[3] + 2) >> 2;
                else if (z & 1) v = F3(L[y + (x >> 1)], L[y + (x >> 1) + 1], L[y + (x >> 1) + 2]);
                else v = F2(L[y + (x >> 1)], L[y + (x >> 1) + 1]);
                break;
            }
            }
            dst[y * 4 + x] = (uint8_t)v;
        }
#undef F3
#undef F2
}

/* ---- intra 8x8 (High profile): reference sample filtering 8.3.2.2.1 + the nine predictors
 * 8.3.2.2.2-8.3.2.2.10.  avail as for 4x4 blocks; TR missing -> p[7,-1] replicated. */
void b2o_pred8x8l(int mode, const uint8_t *p, int pitch, int avail, uint8_t dst[64])
{
    int Tb[17], Lb[9];
    int *T = Tb + 1, *L = Lb + 1;                 /* T[-1] = L[-1] = filtered top-left */
    int t[16], l[8], m = 128;
    const uint8_t *top = p - pitch;
    const int hasT = (avail & B2O_AV_T) != 0, hasL = (avail & B2O_AV_L) != 0, hasTL = (avail & B2O_AV_TL) != 0;
    for (int i = 0; i < 8; i++) { t[i] = hasT ? top[i] : 128; l[i] = hasL ? p[i * pitch - 1] : 128; }
    for (int i = 8; i < 16; i++) t[i] = (avail & B2O_AV_TR) ? top[i] : t[7];
    if (hasTL) m = top[-1];
    /* filtered samples */
    for (int i = 0; i < 16; i++) T[i] = 128;
    for (int i = 0; i < 8; i++) L[i] = 128;
    int M = 128;
    if (hasT) {
        T[0] = hasTL ? (m + 2 * t[0] + t[1] + 2) >> 2 : (3 * t[0] + t[1] + 2) >> 2;
        for (int i = 1; i < 15; i++) T[i] = (t[i - 1] + 2 * t[i] + t[i + 1] + 2) >> 2;
        T[15] = (t[14] + 3 * t[15] + 2) >> 2;
    }
    if (hasTL) {
        if (!hasT) M = (3 * m + l[0] + 2) >> 2;            /* hasL is implied when the corner exists without the top */
        else if (!hasL) M = (3 * m + t[0] + 2) >> 2;
        else M = (t[0] + 2 * m + l[0] + 2) >> 2;
    }
    if (hasL) {
        L[0] = hasTL ? (m + 2 * l[0] + l[1] + 2) >> 2 : (3 * l[0] + l[1] + 2) >> 2;
        for (int i = 1; i < 7; i++) L[i] = (l[i - 1] + 2 * l[i] + l[i + 1] + 2) >> 2;
        L[7] = (l[6] + 3 * l[7] + 2) >> 2;
    }
    T[-1] = L[-1] = M;
#define F3(a, b, c) (((a) + 2 * (b) + (c) + 2) >> 2)
#define F2(a, b) (((a) + (b) + 1) >> 1)
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            int v;
            switch (mode) {
            case B2_I4_V: v = T[x]; break;
            case B2_I4_H: v = L[y]; break;
            case B2_I4_DC: {
                int s = 0;
                if (hasT) for (int i = 0; i < 8; i++) s += T[i];
                if (hasL) for (int i = 0; i < 8; i++) s += L[i];
                v = (hasT && hasL) ? (s + 8) >> 4 : (hasT || hasL) ? (s + 4) >> 3 : 128;
                break;
            }
            case B2_I4_DDL:
                v = (x == 7 && y == 7) ? (T[14] + 3 * T[15] + 2) >> 2 : F3(T[x + y], T[x + y + 1], T[x + y + 2]);
                break;
            case B2_I4_DDR:
                if (x > y) v = F3(T[x - y - 2], T[x - y - 1], T[x - y]);
                else if (x < y) v = F3(L[y - x - 2], L[y - x - 1], L[y - x]);
                else v = F3(T[0], M, L[0]);
                break;
            case B2_I4_VR: {
                int z = 2 * x - y, k = x - (y >> 1);
                if (z >= 0 && !(z & 1)) v = F2(T[k - 1], T[k]);
                else if (z >= 0) v = F3(T[k - 2], T[k - 1], T[k]);
                else if (z == -1) v = F3(L[0], M, T[0]);
                else v = F3(L[y - 2 * x - 1], L[y - 2 * x - 2], L[y - 2 * x - 3]);
                break;
            }
            case B2_I4_HD: {
                int z = 2 * y - x, k = y - (x >> 1);
                if (z >= 0 && !(z & 1)) v = F2(L[k - 1], L[k]);
                else if (z >= 0) v = F3(L[k - 2], L[k - 1], L[k]);
                else if (z == -1) v = F3(L[0], M, T[0]);
                else v = F3(T[x - 2 * y - 1], T[x - 2 * y - 2], T[x - 2 * y - 3]);
                break;
            }
            case B2_I4_VL: {
                int k = x + (y >> 1);
                v = (y & 1) ? F3(T[k], T[k + 1], T[k + 2]) : F2(T[k], T[k + 1]);
                break;
            }
            default: {                                   /* HU */
                int z = x + 2 * y, k = y + (x >> 1);
                if (z > 13) v = L[7];
                else if (z == 13) v = (L[6] + 3 * L[7] + 2) >> 2;
                else if (z & 1) v = F3(L[k], L[k + 1], L[k + 2]);
                else v = F2(L[k], L[k + 1]);
                break;
            }
            }
            dst[y * 8 + x] = (uint8_t)v;
        }
#undef F3
#undef F2
}

/* availability of 8x8 block q (raster 2x2) inside an MB with MB-level availability `mba` */
int b2o_blk8_avail(int q, int mba)
{
    int qx = q & 1, qy = q >> 1, a = 0;
    if (qx || (mba & B2O_AV_L)) a |= B2O_AV_L;
    if (qy || (mba & B2O_AV_T)) a |= B2O_AV_T;
    if ((qx && qy) || (qx && !qy && (mba & B2O_AV_T)) || (!qx && qy && (mba & B2O_AV_L)) || (!qx && !qy && (mba & B2O_AV_TL)))
        a |= B2O_AV_TL;
    if (q == 0 ? (mba & B2O_AV_T) : q == 1 ? (mba & B2O_AV_TR) : q == 2) a |= B2O_AV_TR;
    return a;
}

/* neighbour availability of MB (mbx,mby): one slice per frame, so purely geometric */
int b2o_mb_avail(int mbx, int mby, int mbw)
{
    int a = 0;
    if (mbx > 0) a |= B2O_AV_L;
    if (mby > 0) a |= B2O_AV_T;
    if (mbx > 0 && mby > 0) a |= B2O_AV_TL;
    if (mby > 0 && mbx < mbw - 1) a |= B2O_AV_TR;
    return a;
}

/* availability of 4x4 block b (z order) inside an MB with MB-level availability `mba` */
int b2o_blk_avail(int b, int mba)
{
    int bx = b2o_blk_x[b], by = b2o_blk_y[b], a = 0;
    if (bx > 0 || (mba & B2O_AV_L)) a |= B2O_AV_L;
    if (by > 0 || (mba & B2O_AV_T)) a |= B2O_AV_T;
    if ((bx > 0 && by > 0) || (bx > 0 && (mba & B2O_AV_T)) || (by > 0 && (mba & B2O_AV_L)) ||
        (bx == 0 && by == 0 && (mba & B2O_AV_TL)))
        a |= B2O_AV_TL;
    /* top-right: decoded-before rule */
    if (by == 0) {
        if (bx < 3 ? (mba & B2O_AV_T) : (mba & B2O_AV_TR)) a |= B2O_AV_TR;
    } else if (bx < 3 && b != 3 && b != 11 && b != 7 && b != 13 && b != 15) {
        a |= B2O_AV_TR;
    }
    return a;
}

static const uint8_t ue_bits4[4] = {1, 3, 3, 5};

void b2o_intra_analyse(const b2o_frame_t *cur, int lambda, b2_mbinfo_t *info, uint32_t *cost_i16, uint32_t *cost_i4,
                       uint32_t *cost_i8)
{
    for (int mby = 0; mby < cur->mbh; mby++)
        for (int mbx = 0; mbx < cur->mbw; mbx++) {
            int i = mby * cur->mbw + mbx;
            int mba = b2o_mb_avail(mbx, mby, cur->mbw);
            const uint8_t *sy = cur->y + (size_t)(mby * 16) * cur->pitch + mbx * 16;
            /* I16x16 */
            uint32_t best16 = 0xffffffffu; int m16 = B2_I16_DC;
            for (int m = 0; m < 4; m++) {
                if (!b2o_i16_mode_ok(m, mba)) continue;
                uint8_t pred[256];
                b2o_pred16x16(m, sy, cur->pitch, mba, pred);
                uint32_t c = b2o_satd16x16(sy, cur->pitch, pred, 16) + (uint32_t)lambda * ue_bits4[m];
                if (c < best16) { best16 = c; m16 = m; }
            }
            /* I4x4 */
            uint32_t sum4 = (uint32_t)lambda * 24;
            for (int b = 0; b < 16; b++) {
                const uint8_t *sb = sy + b2o_blk_y[b] * 4 * cur->pitch + b2o_blk_x[b] * 4;
                int ba = b2o_blk_avail(b, mba);
                uint32_t best = 0xffffffffu; int bm = B2_I4_DC;
                for (int m = 0; m < 9; m++) {
                    if (!b2o_i4_mode_ok(m, ba)) continue;
                    uint8_t pred[16];
                    b2o_pred4x4(m, sb, cur->pitch, ba, pred);
                    uint32_t c = b2o_satd4x4(sb, cur->pitch, pred, 4) + (uint32_t)lambda * (m == B2_I4_DC ? 1 : 4);
                    if (c < best) { best = c; bm = m; }
                }
                info[i].i4_mode[b] = (uint8_t)bm;
                sum4 += best;
            }
            /* I8x8 (only with the 8x8 transform): SA8D (8x8 Hadamard) per block like x264's i8x8 analysis, same bit model as I4x4 */
            if (cost_i8) {
                uint32_t sum8 = (uint32_t)lambda * 8;
                unsigned modes = 0;
                for (int q = 0; q < 4; q++) {
                    const uint8_t *sb = sy + (q >> 1) * 8 * cur->pitch + (q & 1) * 8;
                    int ba = b2o_blk8_avail(q, mba);
                    uint32_t best = 0xffffffffu; int bm = B2_I4_DC;
                    for (int m = 0; m < 9; m++) {
                        if (!b2o_i4_mode_ok(m, ba)) continue;
                        uint8_t pred[64];
                        b2o_pred8x8l(m, sb, cur->pitch, ba, pred);
                        uint32_t c = b2o_sa8d8x8(sb, cur->pitch, pred, 8) + (uint32_t)lambda * (m == B2_I4_DC ? 1 : 4);
                        if (c < best) { best = c; bm = m; }
                    }
                    modes |= (unsigned)bm << (4 * q);
                    sum8 += best;
                }
                info[i].i8_modes = (uint16_t)modes;
                cost_i8[i] = sum8;
            }
            /* chroma 8x8 */
            const uint8_t *su = cur->u + (size_t)(mby * 8) * cur->pitchc + mbx * 8;
            const uint8_t *sv = cur->v + (size_t)(mby * 8) * cur->pitchc + mbx * 8;
            uint32_t bestc = 0xffffffffu; int mc = B2_IC_DC;
            for (int m = 0; m < 4; m++) {
                int ok = m == B2_IC_DC ? 1 : m == B2_IC_H ? (mba & B2O_AV_L) != 0 : m == B2_IC_V ? (mba & B2O_AV_T) != 0
                         : (mba & 7) == 7;
                if (!ok) continue;
                uint8_t pu[64], pv[64];
                b2o_pred8x8c(m, su, cur->pitchc, mba, pu);
                b2o_pred8x8c(m, sv, cur->pitchc, mba, pv);
                uint32_t c = b2o_satd8x8(su, cur->pitchc, pu, 8) + b2o_satd8x8(sv, cur->pitchc, pv, 8) +
                             (uint32_t)lambda * ue_bits4[m];
                if (c < bestc) { bestc = c; mc = m; }
            }
            info[i].i16_mode = (uint8_t)m16;
            info[i].chroma_mode = (uint8_t)mc;
            cost_i16[i] = best16;
            cost_i4[i] = sum4;
        }
}
