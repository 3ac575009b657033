/* b2o_me.c -- ORACLE (test infrastructure only; see b2o.h).
 * Motion estimation as BASELINE.json's north_star fixes it: exhaustive full-pel SAD search
 * then half/quarter-pel SATD refinement, per 16x16 macroblock.  In the reference this work
 * happens inside x264_encoder_encode (av_encode.c:970); libx264 is absent, so scan order,
 * tie-break and cost model are the frozen definitions of SURVEY.md Appendix A ("E, ours"). */
#include <stdlib.h>
#include "b2o.h"

void b2o_me_fullpel_mb(const b2o_frame_t *cur, const b2o_frame_t *ref, int R, int mbx, int mby,
                       b2_mv_t pmv, int lambda, b2_mv_t *mv_out, uint32_t *cost_out)
{
    const uint8_t *c = cur->y + (size_t)(mby * 16) * cur->pitch + mbx * 16;
    uint32_t best = 0xffffffffu;
    int bdx = 0, bdy = 0;
    for (int dy = -R; dy <= R; dy++)            /* scan index is dy-major, dx ascending */
        for (int dx = -R; dx <= R; dx++) {
            const uint8_t *r = ref->y + (ptrdiff_t)(mby * 16 + dy) * ref->pitch + mbx * 16 + dx;
            uint32_t cost = b2o_sad16x16(c, cur->pitch, r, ref->pitch)
                          + (uint32_t)lambda * (uint32_t)(b2o_mvbits(4 * dx - pmv.x) + b2o_mvbits(4 * dy - pmv.y));
            if (cost < best) { best = cost; bdx = dx; bdy = dy; }   /* strict <: lowest index wins ties */
        }
    mv_out->x = (int16_t)bdx; mv_out->y = (int16_t)bdy;
    *cost_out = best;
}

void b2o_me_fullpel(const b2o_frame_t *cur, const b2o_frame_t *ref, int R,
                    const b2_mv_t *pmv, int lambda, b2_mv_t *mv_out, uint32_t *cost_out)
{
    for (int mby = 0; mby < cur->mbh; mby++)
        for (int mbx = 0; mbx < cur->mbw; mbx++) {
            int i = mby * cur->mbw + mbx;
            b2_mv_t p = {0, 0};
            if (pmv) p = pmv[i];
            b2o_me_fullpel_mb(cur, ref, R, mbx, mby, p, lambda, &mv_out[i], &cost_out[i]);
        }
}

/* centre first, then the 8 neighbours in raster order */
const int8_t b2o_subpel_offsets[9][2] = {
    {0, 0}, {-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}
};

static uint32_t subpel_cost(const b2o_frame_t *cur, const b2o_frame_t *ref, int mbx, int mby,
                            int mvx, int mvy, b2_mv_t pmv, int lambda)
{
    uint8_t pred[256];
    b2o_mc_luma(ref->y, ref->pitch, mbx * 16, mby * 16, mvx, mvy, 16, 16, pred, 16);
    return b2o_satd16x16(cur->y + (size_t)(mby * 16) * cur->pitch + mbx * 16, cur->pitch, pred, 16)
         + (uint32_t)lambda * (uint32_t)(b2o_mvbits(mvx - pmv.x) + b2o_mvbits(mvy - pmv.y));
}

void b2o_me_subpel(const b2o_frame_t *cur, const b2o_frame_t *ref, const b2_mv_t *mv_full,
                   const b2_mv_t *pmv, int lambda, b2_mv_t *mv_out, uint32_t *cost_out)
{
    for (int mby = 0; mby < cur->mbh; mby++)
        for (int mbx = 0; mbx < cur->mbw; mbx++) {
            int i = mby * cur->mbw + mbx;
            b2_mv_t p = {0, 0};
            if (pmv) p = pmv[i];
            int cx = mv_full[i].x * 4, cy = mv_full[i].y * 4;
            uint32_t best = 0xffffffffu; int bx = cx, by = cy;
            for (int k = 0; k < 9; k++) {                    /* half-pel: step 2 quarter units */
                int mx = cx + 2 * b2o_subpel_offsets[k][0], my = cy + 2 * b2o_subpel_offsets[k][1];
                uint32_t c = subpel_cost(cur, ref, mbx, mby, mx, my, p, lambda);
                if (c < best) { best = c; bx = mx; by = my; }
            }
            cx = bx; cy = by;
            for (int k = 1; k < 9; k++) {                    /* quarter-pel around the half-pel winner */
                int mx = cx + b2o_subpel_offsets[k][0], my = cy + b2o_subpel_offsets[k][1];
                uint32_t c = subpel_cost(cur, ref, mbx, mby, mx, my, p, lambda);
                if (c < best) { best = c; bx = mx; by = my; }
            }
            mv_out[i].x = (int16_t)bx; mv_out[i].y = (int16_t)by;
            cost_out[i] = best;
        }
}

/* ---- sub-pel refinement with inter partitions (16x8, 8x16, 8x8) -----------------------------
 * Same candidate pattern as b2o_me_subpel, evaluated per 8x8 quadrant:
 *   stage 1: the 9 half-pel candidates; for every shape part P (16x16 | 16x8 top,bottom | 8x16 left,right | four 8x8)
 *            cost_P[k] = sum of its quadrants' SATD + lambda*mvbits(mv_k - pmv), first minimum wins;
 *   shape  : 16x16, 16x8, 8x16, 8x8 cost the sum of their parts + lambda * {0, 2, 2, 8} (mb_type / sub_mb_type bits
 *            beyond P_L0_16x16), first minimum wins;
 *   stage 2: every part of the chosen shape tries the 8 quarter-pel neighbours of ITS half-pel winner.
 * All parts therefore stay within +-3 quarter-pels of the 16x16 full-pel vector (a local partition refinement; the
 * full-pel search itself is per macroblock, BASELINE.json north_star).  For shape 16x16 the result equals b2o_me_subpel. */
static const uint8_t part_quads[4][4] = {{0xf, 0, 0, 0}, {0x3, 0xc, 0, 0}, {0x5, 0xa, 0, 0}, {0x1, 0x2, 0x4, 0x8}};   /* quadrant masks per part */
static const uint8_t part_count[4] = {1, 2, 2, 4};
static const uint8_t part_bits[4] = {0, 2, 2, 8};

static void quad_satd(const b2o_frame_t *cur, const b2o_frame_t *ref, int mbx, int mby, int mvx, int mvy, uint32_t out[4])
{
    uint8_t pred[256];
    const uint8_t *c = cur->y + (size_t)(mby * 16) * cur->pitch + mbx * 16;
    b2o_mc_luma(ref->y, ref->pitch, mbx * 16, mby * 16, mvx, mvy, 16, 16, pred, 16);
    for (int q = 0; q < 4; q++)
        out[q] = b2o_satd8x8(c + (q >> 1) * 8 * cur->pitch + (q & 1) * 8, cur->pitch, pred + (q >> 1) * 8 * 16 + (q & 1) * 8, 16);
}

void b2o_me_subpel_part(const b2o_frame_t *cur, const b2o_frame_t *ref, const b2_mv_t *mv_full, const b2_mv_t *pmv, int lambda,
                        uint8_t *part_out, b2_mv_t (*mv_out)[4], uint32_t *cost_out)
{
    for (int mby = 0; mby < cur->mbh; mby++)
        for (int mbx = 0; mbx < cur->mbw; mbx++) {
            const int i = mby * cur->mbw + mbx;
            b2_mv_t p = {0, 0};
            if (pmv) p = pmv[i];
            const int cx = mv_full[i].x * 4, cy = mv_full[i].y * 4;
            uint32_t sq[9][4];
            for (int k = 0; k < 9; k++)
                quad_satd(cur, ref, mbx, mby, cx + 2 * b2o_subpel_offsets[k][0], cy + 2 * b2o_subpel_offsets[k][1], sq[k]);
            /* stage 1: best half-pel candidate of every part of every shape */
            uint32_t pc[4][4]; int pk[4][4];
            uint32_t shape_cost[4]; int shape = 0;
            for (int sh = 0; sh < 4; sh++) {
                shape_cost[sh] = (uint32_t)lambda * part_bits[sh];
                for (int a = 0; a < part_count[sh]; a++) {
                    uint32_t best = 0xffffffffu; int bk = 0;
                    for (int k = 0; k < 9; k++) {
                        const int mx = cx + 2 * b2o_subpel_offsets[k][0], my = cy + 2 * b2o_subpel_offsets[k][1];
                        uint32_t c = (uint32_t)lambda * (uint32_t)(b2o_mvbits(mx - p.x) + b2o_mvbits(my - p.y));
                        for (int q = 0; q < 4; q++)
                            if (part_quads[sh][a] & (1 << q)) c += sq[k][q];
                        if (c < best) { best = c; bk = k; }
                    }
                    pc[sh][a] = best; pk[sh][a] = bk;
                    shape_cost[sh] += best;
                }
                if (shape_cost[sh] < shape_cost[shape]) shape = sh;
            }
            /* stage 2: quarter-pel neighbours of each part's winner */
            uint32_t total = (uint32_t)lambda * part_bits[shape];
            for (int a = 0; a < part_count[shape]; a++) {
                const int hx = cx + 2 * b2o_subpel_offsets[pk[shape][a]][0], hy = cy + 2 * b2o_subpel_offsets[pk[shape][a]][1];
                uint32_t best = pc[shape][a]; int bx = hx, by = hy;
                for (int k = 1; k < 9; k++) {
                    const int mx = hx + b2o_subpel_offsets[k][0], my = hy + b2o_subpel_offsets[k][1];
                    uint32_t s4[4];
                    quad_satd(cur, ref, mbx, mby, mx, my, s4);
                    uint32_t c = (uint32_t)lambda * (uint32_t)(b2o_mvbits(mx - p.x) + b2o_mvbits(my - p.y));
                    for (int q = 0; q < 4; q++)
                        if (part_quads[shape][a] & (1 << q)) c += s4[q];
                    if (c < best) { best = c; bx = mx; by = my; }
                }
                total += best;
                for (int q = 0; q < 4; q++)
                    if (part_quads[shape][a] & (1 << q)) { mv_out[i][q].x = (int16_t)bx; mv_out[i][q].y = (int16_t)by; }
            }
            part_out[i] = (uint8_t)shape;
            cost_out[i] = total;
        }
}

/* ---- wide partition search (row N1, partitions = 2): every shape part gets its own exhaustive full-pel search ----------
 * One pass over the (2R+1)^2 candidates yields the SAD of each 8x8 quadrant; the nine shape parts (16x16 | 16x8 top,bottom |
 * 8x16 left,right | four 8x8) are sums of quadrants, cost_P = SAD_P + lambda*mvbits(4d - pmv), lexicographic min of
 * (cost, scan index) per part exactly like b2o_me_fullpel.  The shape is chosen on these full-pel costs (+ lambda*{0,2,2,8});
 * each part of the chosen shape is then refined like b2o_me_subpel, with SATD over its own area.  Shape 16x16 reproduces
 * b2o_me_fullpel + b2o_me_subpel. */
void b2o_me_fullpel_parts_mb(const b2o_frame_t *cur, const b2o_frame_t *ref, int R, int mbx, int mby, b2_mv_t pmv, int lambda,
                             b2_mv_t mv9[9], uint32_t cost9[9])
{
    static const uint8_t mask9[9] = {0xf, 0x3, 0xc, 0x5, 0xa, 0x1, 0x2, 0x4, 0x8};
    const uint8_t *c = cur->y + (size_t)(mby * 16) * cur->pitch + mbx * 16;
    for (int p = 0; p < 9; p++) { cost9[p] = 0xffffffffu; mv9[p].x = mv9[p].y = 0; }
    for (int dy = -R; dy <= R; dy++)
        for (int dx = -R; dx <= R; dx++) {
            const uint8_t *r = ref->y + (ptrdiff_t)(mby * 16 + dy) * ref->pitch + mbx * 16 + dx;
            uint32_t q[4] = {0, 0, 0, 0};
            for (int y = 0; y < 16; y++)
                for (int x = 0; x < 16; x++) q[(x >> 3) | ((y >> 3) << 1)] += (uint32_t)abs(c[y * cur->pitch + x] - r[y * ref->pitch + x]);
            const uint32_t mvc = (uint32_t)lambda * (uint32_t)(b2o_mvbits(4 * dx - pmv.x) + b2o_mvbits(4 * dy - pmv.y));
            for (int p = 0; p < 9; p++) {
                uint32_t s = mvc;
                for (int k = 0; k < 4; k++)
                    if (mask9[p] & (1 << k)) s += q[k];
                if (s < cost9[p]) { cost9[p] = s; mv9[p].x = (int16_t)dx; mv9[p].y = (int16_t)dy; }
            }
        }
}

/* SATD of the w x h block at (x0,y0) of the macroblock against the prediction displaced by quarter-pel (mvx,mvy) */
static uint32_t part_satd(const b2o_frame_t *cur, const b2o_frame_t *ref, int px, int py, int w, int h, int mvx, int mvy)
{
    uint8_t pred[256];
    uint32_t s = 0;
    b2o_mc_luma(ref->y, ref->pitch, px, py, mvx, mvy, w, h, pred, 16);
    for (int y = 0; y < h; y += 4)
        for (int x = 0; x < w; x += 4)
            s += b2o_satd4x4(cur->y + (size_t)(py + y) * cur->pitch + px + x, cur->pitch, pred + y * 16 + x, 16);
    return s;
}

void b2o_me_parts_wide(const b2o_frame_t *cur, const b2o_frame_t *ref, int R, const b2_mv_t *pmv, int lambda, int subpel,
                       b2_mv_t *mv16_full, uint32_t *cost16_full, uint8_t *part_out, b2_mv_t (*mv_out)[4], uint32_t *cost_out)
{
    static const uint8_t first[4] = {0, 1, 3, 5}, count[4] = {1, 2, 2, 4}, bits[4] = {0, 2, 2, 8};
    /* geometry of part p in pixels: x, y, w, h */
    static const uint8_t geo[9][4] = {{0, 0, 16, 16}, {0, 0, 16, 8}, {0, 8, 16, 8}, {0, 0, 8, 16}, {8, 0, 8, 16},
                                      {0, 0, 8, 8},   {8, 0, 8, 8},  {0, 8, 8, 8},  {8, 8, 8, 8}};
    for (int mby = 0; mby < cur->mbh; mby++)
        for (int mbx = 0; mbx < cur->mbw; mbx++) {
            const int i = mby * cur->mbw + mbx;
            b2_mv_t p = {0, 0};
            if (pmv) p = pmv[i];
            b2_mv_t mv9[9]; uint32_t c9[9];
            b2o_me_fullpel_parts_mb(cur, ref, R, mbx, mby, p, lambda, mv9, c9);
            mv16_full[i] = mv9[0]; cost16_full[i] = c9[0];
            int shape = 0; uint32_t bc = 0xffffffffu;
            for (int sh = 0; sh < 4; sh++) {
                uint32_t c = (uint32_t)lambda * bits[sh];
                for (int a = 0; a < count[sh]; a++) c += c9[first[sh] + a];
                if (c < bc) { bc = c; shape = sh; }
            }
            if (!subpel) shape = 0;                          /* partitions need the sub-pel stage (SATD-domain costs) */
            uint32_t total = (uint32_t)lambda * bits[shape];
            for (int a = 0; a < count[shape]; a++) {
                const int pp = first[shape] + a;
                const int px = mbx * 16 + geo[pp][0], py = mby * 16 + geo[pp][1], w = geo[pp][2], h = geo[pp][3];
                int cx = mv9[pp].x * 4, cy = mv9[pp].y * 4;
                uint32_t best = 0xffffffffu; int bx = cx, by = cy;
                for (int k = 0; k < (subpel ? 9 : 1); k++) {
                    const int mx = cx + 2 * b2o_subpel_offsets[k][0], my = cy + 2 * b2o_subpel_offsets[k][1];
                    uint32_t c = part_satd(cur, ref, px, py, w, h, mx, my) + (uint32_t)lambda * (uint32_t)(b2o_mvbits(mx - p.x) + b2o_mvbits(my - p.y));
                    if (c < best) { best = c; bx = mx; by = my; }
                }
                cx = bx; cy = by;
                for (int k = 1; k < (subpel ? 9 : 1); k++) {
                    const int mx = cx + b2o_subpel_offsets[k][0], my = cy + b2o_subpel_offsets[k][1];
                    uint32_t c = part_satd(cur, ref, px, py, w, h, mx, my) + (uint32_t)lambda * (uint32_t)(b2o_mvbits(mx - p.x) + b2o_mvbits(my - p.y));
                    if (c < best) { best = c; bx = mx; by = my; }
                }
                total += best;
                for (int q = 0; q < 4; q++) {
                    const int qx = (q & 1) * 8, qy = (q >> 1) * 8;
                    if (qx >= geo[pp][0] && qx < geo[pp][0] + w && qy >= geo[pp][1] && qy < geo[pp][1] + h) { mv_out[i][q].x = (int16_t)bx; mv_out[i][q].y = (int16_t)by; }
                }
            }
            part_out[i] = (uint8_t)shape;
            cost_out[i] = total;
        }
}
