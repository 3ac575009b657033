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
