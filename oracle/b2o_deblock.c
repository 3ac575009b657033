/* b2o_deblock.c -- ORACLE (test infrastructure only; see b2o.h).
 * In-loop deblocking filter, ITU-T H.264 8.7 (normative; pinned by the libavcodec decoder drift test), restricted to
 * what the stage produces: frame macroblocks, 4x4 or 8x8 transform (8x8: only the 8-pel edges are transform edges),
 * one reference frame, constant QP, slice-level filter offsets (8.7.2.2: indexA = qPav + 2*slice_alpha_c0_offset_div2,
 * indexB = qPav + 2*slice_beta_offset_div2, both clipped to 0..51).
 * In the reference this is part of x264_encoder_encode (av_encode.c:970; x264 enables it by default and the reference's
 * default tune "film" (av_encode.c:103) sets the offsets to -1:-1).  SURVEY.md 8f row N2. */
#include <stdlib.h>
#include "b2o.h"

static const uint8_t alpha_tab[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13,
                                      15, 17, 20, 22, 25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144,
                                      162, 182, 203, 226, 255, 255};
static const uint8_t beta_tab[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4,
                                     6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18};
static const uint8_t tc0_tab[52][3] = {
    {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0},
    {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 1, 1},
    {0, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 2, 3}, {1, 2, 3},
    {2, 2, 3}, {2, 2, 4}, {2, 3, 4}, {2, 3, 4}, {3, 3, 5}, {3, 4, 6}, {3, 4, 6}, {4, 5, 7}, {4, 5, 8}, {4, 6, 9}, {5, 7, 10},
    {6, 8, 11}, {6, 8, 13}, {7, 10, 14}, {8, 11, 16}, {9, 12, 18}, {10, 13, 20}, {11, 15, 23}, {13, 17, 25}};

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* filter one line of samples across an edge; pix -> q0, `step` = distance between successive samples across the edge */
static void filter_line(uint8_t *pix, int step, int bs, int ia, int ib, int chroma)
{
    if (!bs) return;
    const int alpha = alpha_tab[ia], beta = beta_tab[ib];
    const int p0 = pix[-step], p1 = pix[-2 * step], q0 = pix[0], q1 = pix[step];
    if (!(abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta)) return;
    if (chroma) {
        if (bs < 4) {
            const int tc = tc0_tab[ia][bs - 1] + 1;
            const int d = clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
            pix[-step] = (uint8_t)clip255(p0 + d); pix[0] = (uint8_t)clip255(q0 - d);
        } else {
            pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
            pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
        }
        return;
    }
    const int p2 = pix[-3 * step], q2 = pix[2 * step];
    const int ap = abs(p2 - p0), aq = abs(q2 - q0);
    if (bs < 4) {
        const int tc0 = tc0_tab[ia][bs - 1];
        const int tc = tc0 + (ap < beta) + (aq < beta);
        const int d = clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
        pix[-step] = (uint8_t)clip255(p0 + d); pix[0] = (uint8_t)clip255(q0 - d);
        if (ap < beta) pix[-2 * step] = (uint8_t)(p1 + clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1));
        if (aq < beta) pix[step] = (uint8_t)(q1 + clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1));
    } else {
        const int p3 = pix[-4 * step], q3 = pix[3 * step];
        const int strong = abs(p0 - q0) < ((alpha >> 2) + 2);
        if (ap < beta && strong) {
            pix[-step] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            pix[-2 * step] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            pix[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else {
            pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        }
        if (aq < beta && strong) {
            pix[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            pix[step] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
            pix[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else {
            pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
        }
    }
}

/* z-order index of the 4x4 block at (bx,by) */
static inline int zidx(int bx, int by) { return (bx & 1) | ((by & 1) << 1) | ((bx >> 1) << 2) | ((by >> 1) << 3); }

/* does the transform block that contains 4x4 block (bx,by) hold non-zero coefficients?  With the 8x8 transform
 * that is the whole 8x8 block (nnz_mask bits 4q..4q+3 are its interleaved quarters) */
static inline int blk_coded(const b2_mbinfo_t *m, int bx, int by)
{
    if (m->transform8x8) return ((m->nnz_mask >> (4 * ((bx >> 1) | ((by >> 1) << 1)))) & 15u) != 0;
    return (m->nnz_mask >> zidx(bx, by)) & 1u;
}

/* boundary strength between 4x4 block (pbx,pby) of MB `mp` and block (qbx,qby) of MB `mq` (8.7.2.1) */
static int bs_of(const b2_mbinfo_t *mp, int pbx, int pby, const b2_mbinfo_t *mq, int qbx, int qby, int mb_edge)
{
    if (mp->mb_type != B2_MB_P16x16 || mq->mb_type != B2_MB_P16x16) return mb_edge ? 4 : 3;
    if (blk_coded(mp, pbx, pby) || blk_coded(mq, qbx, qby)) return 2;
    const b2_mv_t a = b2o_quad_mv(mp, (pbx >> 1) | ((pby >> 1) << 1)), b = b2o_quad_mv(mq, (qbx >> 1) | ((qby >> 1) << 1));
    if (abs(a.x - b.x) >= 4 || abs(a.y - b.y) >= 4) return 1;
    return 0;
}

void b2o_deblock_frame(b2o_frame_t *f, const b2_mbinfo_t *info, int qp, int alpha_off, int beta_off)
{
    const int qpc = b2o_chroma_qp(qp);
    const int ia = clip3(0, 51, qp + 2 * alpha_off), ib = clip3(0, 51, qp + 2 * beta_off);
    const int iac = clip3(0, 51, qpc + 2 * alpha_off), ibc = clip3(0, 51, qpc + 2 * beta_off);
    for (int mby = 0; mby < f->mbh; mby++)
        for (int mbx = 0; mbx < f->mbw; mbx++) {
            const b2_mbinfo_t *mq = &info[mby * f->mbw + mbx];
            uint8_t *y = f->y + (size_t)(mby * 16) * f->pitch + mbx * 16;
            uint8_t *c[2] = {f->u + (size_t)(mby * 8) * f->pitchc + mbx * 8, f->v + (size_t)(mby * 8) * f->pitchc + mbx * 8};
            int bs[4][4];
            /* vertical edges (left to right) */
            for (int e = 0; e < 4; e++) {
                if ((e == 0 && mbx == 0) || ((e & 1) && mq->transform8x8)) { for (int k = 0; k < 4; k++) bs[e][k] = 0; continue; }
                const b2_mbinfo_t *mp = e == 0 ? mq - 1 : mq;
                for (int k = 0; k < 4; k++) bs[e][k] = bs_of(mp, e == 0 ? 3 : e - 1, k, mq, e, k, e == 0);
                for (int r = 0; r < 16; r++) filter_line(y + (size_t)r * f->pitch + 4 * e, 1, bs[e][r >> 2], ia, ib, 0);
            }
            for (int p = 0; p < 2; p++)
                for (int e = 0; e < 2; e++)
                    for (int r = 0; r < 8; r++) filter_line(c[p] + (size_t)r * f->pitchc + 4 * e, 1, bs[2 * e][r >> 1], iac, ibc, 1);
            /* horizontal edges (top to bottom) */
            for (int e = 0; e < 4; e++) {
                if ((e == 0 && mby == 0) || ((e & 1) && mq->transform8x8)) { for (int k = 0; k < 4; k++) bs[e][k] = 0; continue; }
                const b2_mbinfo_t *mp = e == 0 ? mq - f->mbw : mq;
                for (int k = 0; k < 4; k++) bs[e][k] = bs_of(mp, k, e == 0 ? 3 : e - 1, mq, k, e, e == 0);
                for (int x = 0; x < 16; x++) filter_line(y + (size_t)(4 * e) * f->pitch + x, f->pitch, bs[e][x >> 2], ia, ib, 0);
            }
            for (int p = 0; p < 2; p++)
                for (int e = 0; e < 2; e++)
                    for (int x = 0; x < 8; x++) filter_line(c[p] + (size_t)(4 * e) * f->pitchc + x, f->pitchc, bs[2 * e][x >> 1], iac, ibc, 1);
        }
}
