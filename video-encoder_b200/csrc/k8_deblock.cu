// k8_deblock.cu -- K8: in-loop deblocking filter (ITU-T H.264 8.7) on the reconstructed frame.
//
// x264 runs the loop filter by default inside x264_encoder_encode (av_encode.c:970).
// Bit-exact against oracle/b2o_deblock.c and, through the decoder drift test, against libavcodec.
//
// The filter is defined in macroblock raster order and every macroblock reads samples its left, top and
// top-right neighbours have already modified: MB (x, y) may run once (x-1, y) and (x+1, y-1) are done.
//
// Schedule: ROW PIPELINE.  One warp owns one macroblock row and walks it left to right; row r trails row r-1 by two
// macroblocks.  The only cross-warp communication is one progress counter per row: "x+1 macroblocks done", in shared
// memory when producer and consumer row sit in the same CTA (15 of 16 rows), in global memory across CTAs.  The pixels
// travel through L2 (fence.gpu before the flag store, fence + ld.global.cg after the flag load).  The dependency depth is
// the same mbw + 2*mbh steps as the anti-diagonal wavefront, but a step is one warp's macroblock (no cluster-wide barrier:
// 7 us per diagonal before) and the loads are off the chain: a warp prefetches the next macroblock's own samples and
// decision records while it filters the current one, carries the four left columns in shared memory, and only fetches
// the four rows above after the flag.  Plain launch (no cluster to place): a 1080p frame is five half-SM CTAs that slot
// in beside the other stream groups' kernels; waiting warps sleep instead of spinning on the issue slots.
// Per macroblock: the 20x20 luma and two 12x12 chroma neighbourhoods live in a per-warp shared-memory tile, lanes
// 0-15 filter the luma lines (rows for vertical edges, columns for horizontal edges), lanes 16-31 the U and V lines;
// the four edges of a line are filtered by the same lane in order.
// (The anti-diagonal cluster-barrier form is kept behind B2_K8_WAVEFRONT=1 for A/B measurements.)
// Bound: latency of the dependency chain; algorithmic bytes 1.5*W*H read + written.
#include <stdlib.h>
#include "b2_mbcode.cuh"

namespace {

using namespace b2;

// measured: 8 instead of 16 warps per CTA (smaller register footprint beside other groups' K1 CTAs) changes nothing at 1080p
// (3,850 frames/s with deblocking either way) and halves the warps available to 4K's 120-MB diagonals
constexpr int K8_WARPS = 16;
constexpr int LP = 24;      // luma tile pitch: rows -4..15, cols -4..15 -> tile[(y+4)*LP + x + 4]
constexpr int CP = 16;      // chroma tile pitch: rows -2..7, cols -4..7 -> tile[(y+2)*CP + x + 4]

__device__ __constant__ uint8_t c_alpha[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13,
                                               15, 17, 20, 22, 25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144,
                                               162, 182, 203, 226, 255, 255};
__device__ __constant__ uint8_t c_beta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4,
                                              6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18};
__device__ __constant__ uint8_t c_tc0[52][3] = {
    {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0},
    {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 1, 1},
    {0, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 2, 3}, {1, 2, 3},
    {2, 2, 3}, {2, 2, 4}, {2, 3, 4}, {2, 3, 4}, {3, 3, 5}, {3, 4, 6}, {3, 4, 6}, {4, 5, 7}, {4, 5, 8}, {4, 6, 9}, {5, 7, 10},
    {6, 8, 11}, {6, 8, 13}, {7, 10, 14}, {8, 11, 16}, {9, 12, 18}, {10, 13, 20}, {11, 15, 23}, {13, 17, 25}};

struct K8Warp {
    __align__(16) uint8_t y[20 * LP];
    __align__(16) uint8_t c[2][10 * CP];
    int8_t bs[2][4][4];                  // [0 vertical | 1 horizontal][edge][4-sample segment]
};

__device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }

__device__ __forceinline__ int clip3(int lo, int hi, int v) { return min(max(v, lo), hi); }

struct FiltConst { int alpha, beta, tc0[3]; };

// one line of samples across an edge; pix -> q0, step = distance between successive samples across the edge
__device__ __forceinline__ void filter_line(uint8_t *pix, int step, int bs, const FiltConst &fc, bool chroma)
{
    if (bs == 0) return;
    const int p0 = pix[-step], p1 = pix[-2 * step], q0 = pix[0], q1 = pix[step];
    if (!(abs(p0 - q0) < fc.alpha && abs(p1 - p0) < fc.beta && abs(q1 - q0) < fc.beta)) return;
    if (chroma) {
        if (bs < 4) {
            const int tc = fc.tc0[bs - 1] + 1;
            const int d = clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
            pix[-step] = (uint8_t)b2_clip255(p0 + d); pix[0] = (uint8_t)b2_clip255(q0 - d);
        } else {
            pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
            pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
        }
        return;
    }
    const int p2 = pix[-3 * step], q2 = pix[2 * step];
    const int ap = abs(p2 - p0), aq = abs(q2 - q0);
    if (bs < 4) {
        const int tc0 = fc.tc0[bs - 1];
        const int tc = tc0 + (ap < fc.beta) + (aq < fc.beta);
        const int d = clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
        pix[-step] = (uint8_t)b2_clip255(p0 + d); pix[0] = (uint8_t)b2_clip255(q0 - d);
        if (ap < fc.beta) pix[-2 * step] = (uint8_t)(p1 + clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1));
        if (aq < fc.beta) pix[step] = (uint8_t)(q1 + clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1));
    } else {
        const int p3 = pix[-4 * step], q3 = pix[3 * step];
        const bool strong = abs(p0 - q0) < ((fc.alpha >> 2) + 2);
        if (ap < fc.beta && strong) {
            pix[-step] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            pix[-2 * step] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            pix[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else {
            pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        }
        if (aq < fc.beta && strong) {
            pix[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            pix[step] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
            pix[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else {
            pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
        }
    }
}

__device__ __forceinline__ int zidx(int x, int y) { return (x & 1) | ((y & 1) << 1) | ((x >> 1) << 2) | ((y >> 1) << 3); }

// does the transform block containing 4x4 block (bx,by) hold non-zero coefficients?  With the 8x8 transform that is the
// whole 8x8 block (nnz_mask bits 4q..4q+3 = its interleaved quarters)
__device__ __forceinline__ bool blk_coded(const b2_mbinfo_t *m, int bx, int by)
{
    if (m->transform8x8) return ((m->nnz_mask >> (4 * ((bx >> 1) | ((by >> 1) << 1)))) & 15u) != 0;
    return (m->nnz_mask >> zidx(bx, by)) & 1u;
}

// boundary strength (8.7.2.1) between 4x4 block (pbx,pby) of MB p and block (qbx,qby) of MB q
__device__ __forceinline__ int bs_of(const b2_mbinfo_t *mp, int pbx, int pby, const b2_mbinfo_t *mq, int qbx, int qby, bool mb_edge)
{
    if (mp->mb_type != B2_MB_P16x16 || mq->mb_type != B2_MB_P16x16) return mb_edge ? 4 : 3;
    if (blk_coded(mp, pbx, pby) || blk_coded(mq, qbx, qby)) return 2;
    const int pq = (pbx >> 1) | ((pby >> 1) << 1), qq = (qbx >> 1) | ((qby >> 1) << 1);      // 8x8 quadrants: one vector each
    const b2_mv_t a = (pq && mp->part != B2_PART_16x16) ? mp->mv8[pq - 1] : b2_mv_t{mp->mvx, mp->mvy};
    const b2_mv_t b = (qq && mq->part != B2_PART_16x16) ? mq->mv8[qq - 1] : b2_mv_t{mq->mvx, mq->mvy};
    if (abs(a.x - b.x) >= 4 || abs(a.y - b.y) >= 4) return 1;
    return 0;
}

__device__ void k8_mb_task(int lane, K8Warp &ws, uint8_t *const rec[3], int pitch, int pitchc, size_t offy, size_t offc,
                           int mbx, int mby, int mbw, const b2_mbinfo_t *mi, const FiltConst &fy, const FiltConst &fcc)
{
    uint8_t *gy = rec[0] + offy + (size_t)(B2_PAD + mby * 16) * pitch + B2_PAD + mbx * 16;
    // ---- stage the neighbourhood: luma rows -4..15 x cols -4..15 (5 words per row), chroma rows -2..7 x cols -4..7 ----
    for (int i = lane; i < 100; i += 32) {
        const int r = i / 5, w = i - r * 5;
        *(uint32_t *)&ws.y[r * LP + w * 4] = *(const uint32_t *)(gy + (ptrdiff_t)(r - 4) * pitch + (w - 1) * 4);
    }
    for (int i = lane; i < 60; i += 32) {
        const int p = i / 30, k = i - p * 30, r = k / 3, w = k - r * 3;
        const uint8_t *gc = rec[1 + p] + offc + (size_t)(B2_PADC + mby * 8) * pitchc + B2_PADC + mbx * 8;
        *(uint32_t *)&ws.c[p][r * CP + w * 4] = *(const uint32_t *)(gc + (ptrdiff_t)(r - 2) * pitchc + (w - 1) * 4);
    }
    // ---- boundary strengths: lanes 0-15 vertical edges, 16-31 horizontal edges; lane -> (edge e, segment k) ----
    {
        const int dir = lane >> 4, e = (lane >> 2) & 3, k = lane & 3;
        // picture border; with the 8x8 transform only the 8-pel edges are transform-block edges (8.7)
        const bool edge_ok = !(e == 0 && (dir == 0 ? mbx == 0 : mby == 0)) && !((e & 1) && mi->transform8x8);
        int bs = 0;
        if (edge_ok) {
            const b2_mbinfo_t *mp = e == 0 ? (dir == 0 ? mi - 1 : mi - mbw) : mi;
            const int pbx = dir == 0 ? (e == 0 ? 3 : e - 1) : k, pby = dir == 0 ? k : (e == 0 ? 3 : e - 1);
            const int qbx = dir == 0 ? e : k, qby = dir == 0 ? k : e;
            bs = bs_of(mp, pbx, pby, mi, qbx, qby, e == 0);
        }
        ws.bs[dir][e][k] = (int8_t)bs;
    }
    __syncwarp();
    // ---- vertical edges: luma lane = row, chroma lane-16 = (plane, row) ----
    if (lane < 16) {
        uint8_t *row = &ws.y[(lane + 4) * LP + 4];
#pragma unroll
        for (int e = 0; e < 4; e++) filter_line(row + 4 * e, 1, ws.bs[0][e][lane >> 2], fy, false);
    } else {
        const int l = lane - 16, p = l >> 3, r = l & 7;
        uint8_t *row = &ws.c[p][(r + 2) * CP + 4];
#pragma unroll
        for (int e = 0; e < 2; e++) filter_line(row + 4 * e, 1, ws.bs[0][2 * e][r >> 1], fcc, true);
    }
    __syncwarp();
    // ---- horizontal edges: luma lane = column, chroma lane-16 = (plane, column) ----
    if (lane < 16) {
        uint8_t *col = &ws.y[4 * LP + lane + 4];
#pragma unroll
        for (int e = 0; e < 4; e++) filter_line(col + 4 * e * LP, LP, ws.bs[1][e][lane >> 2], fy, false);
    } else {
        const int l = lane - 16, p = l >> 3, x = l & 7;
        uint8_t *col = &ws.c[p][2 * CP + x + 4];
#pragma unroll
        for (int e = 0; e < 2; e++) filter_line(col + 4 * e * CP, CP, ws.bs[1][2 * e][x >> 1], fcc, true);
    }
    __syncwarp();
    // ---- write back what can have changed: own MB, 3 (luma) / 1 (chroma) lines into the top MB, one word into the left MB ----
    for (int i = lane; i < 95; i += 32) {                  // luma rows -3..15 (19 rows) x 5 words
        const int r = i / 5 + 1, w = i - (i / 5) * 5;       // tile row r = y + 4, y = -3..15
        const int y = r - 4;
        if (w == 0 && (mbx == 0 || y < 0)) continue;        // nothing left of the picture; corner block is never modified
        if (y < 0 && mby == 0) continue;
        *(uint32_t *)(gy + (ptrdiff_t)y * pitch + (w - 1) * 4) = *(const uint32_t *)&ws.y[r * LP + w * 4];
    }
    for (int i = lane; i < 54; i += 32) {                  // chroma rows -1..7 (9 rows) x 3 words x 2 planes
        const int p = i / 27, k = i - p * 27, r = k / 3 + 1, w = k - (k / 3) * 3;
        const int y = r - 2;
        if (w == 0 && (mbx == 0 || y < 0)) continue;
        if (y < 0 && mby == 0) continue;
        uint8_t *gc = rec[1 + p] + offc + (size_t)(B2_PADC + mby * 8) * pitchc + B2_PADC + mbx * 8;
        *(uint32_t *)(gc + (ptrdiff_t)y * pitchc + (w - 1) * 4) = *(const uint32_t *)&ws.c[p][r * CP + w * 4];
    }
}

__global__ void __launch_bounds__(K8_WARPS * 32)
k8_deblock_kernel(uint8_t *ry, uint8_t *ru, uint8_t *rv, int pitch, int pitchc, size_t stride_y, size_t stride_c, int mbw, int mbh,
                  int qp, int alpha_off, int beta_off, const b2_mbinfo_t *__restrict__ info)
{
    __shared__ K8Warp s_warp[K8_WARPS];
    const int frame = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncta = (int)cluster_nctarank(), crank = (int)cluster_ctarank();
    const int gwarp = crank * K8_WARPS + warp, nwarps = ncta * K8_WARPS;
    const int ndiag = mbw + 2 * (mbh - 1);
    const b2_mbinfo_t *finfo = info + (size_t)frame * mbw * mbh;
    uint8_t *const rec[3] = {ry, ru, rv};
    const int qpc = chroma_qp(qp);
    FiltConst fy, fcc;
    // 8.7.2.2: indexA = qPav + 2 * slice_alpha_c0_offset_div2, indexB = qPav + 2 * slice_beta_offset_div2, clipped to 0..51
    const int ia = clip3(0, 51, qp + 2 * alpha_off), ib = clip3(0, 51, qp + 2 * beta_off);
    const int iac = clip3(0, 51, qpc + 2 * alpha_off), ibc = clip3(0, 51, qpc + 2 * beta_off);
    fy.alpha = c_alpha[ia]; fy.beta = c_beta[ib]; fcc.alpha = c_alpha[iac]; fcc.beta = c_beta[ibc];
#pragma unroll
    for (int i = 0; i < 3; i++) { fy.tc0[i] = c_tc0[ia][i]; fcc.tc0[i] = c_tc0[iac][i]; }
    for (int d = 0; d < ndiag; d++) {
        const int y_lo = max(0, (d - mbw + 2) >> 1), y_hi = min(mbh - 1, d >> 1);
        for (int t = y_lo + gwarp; t <= y_hi; t += nwarps) {
            const int mby = t, mbx = d - 2 * mby;
            k8_mb_task(lane, s_warp[warp], rec, pitch, pitchc, frame * stride_y, frame * stride_c, mbx, mby, mbw,
                       &finfo[mby * mbw + mbx], fy, fcc);
        }
        if (ncta > 1) cluster_barrier();
        else __syncthreads();
    }
}


// ---- row-pipelined form -------------------------------------------------------------------------------------------
// what the boundary-strength rule needs of one b2_mbinfo_t (48 bytes = three 16-byte words: {mv, type.., i4_mode[0..7]},
// {i4_mode[8..15], cost, nnz_mask}, {mv8[3], part | transform8x8 << 8 | i8_modes << 16})
struct MbLite { uint32_t mv, typ, nnz, m8a, m8b, m8c, pt; };

__device__ __forceinline__ MbLite load_lite(const b2_mbinfo_t *m)
{
    const uint4 *q = reinterpret_cast<const uint4 *>(m);
    const uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    MbLite l;
    l.mv = a.x; l.typ = a.y & 255u; l.nnz = b.w; l.m8a = c.x; l.m8b = c.y; l.m8c = c.z; l.pt = c.w;
    return l;
}
__device__ __forceinline__ bool lite_coded(const MbLite &m, int bx, int by)
{
    if ((m.pt >> 8) & 255u) return ((m.nnz >> (4 * ((bx >> 1) | ((by >> 1) << 1)))) & 15u) != 0;
    return (m.nnz >> zidx(bx, by)) & 1u;
}
__device__ __forceinline__ uint32_t lite_mv(const MbLite &m, int q)
{
    if (q == 0 || (m.pt & 255u) == B2_PART_16x16) return m.mv;
    return q == 1 ? m.m8a : (q == 2 ? m.m8b : m.m8c);
}
__device__ __forceinline__ int lite_bs(const MbLite &mp, int pbx, int pby, const MbLite &mq, int qbx, int qby, bool mb_edge)
{
    if (mp.typ != B2_MB_P16x16 || mq.typ != B2_MB_P16x16) return mb_edge ? 4 : 3;
    if (lite_coded(mp, pbx, pby) || lite_coded(mq, qbx, qby)) return 2;
    const uint32_t a = lite_mv(mp, (pbx >> 1) | ((pby >> 1) << 1)), b = lite_mv(mq, (qbx >> 1) | ((qby >> 1) << 1));
    const int ax = (int16_t)(a & 0xffffu), ay = (int16_t)(a >> 16), bx = (int16_t)(b & 0xffffu), by = (int16_t)(b >> 16);
    return (abs(ax - bx) >= 4 || abs(ay - by) >= 4) ? 1 : 0;
}

// cross-CTA progress flags live in global memory (cleared by the launcher ahead of every launch)
__device__ __forceinline__ void st_flag_global(int *f, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_flag_global(const int *f) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory"); return v; }

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2)
k8_deblock_rows_kernel(uint8_t *ry, uint8_t *ru, uint8_t *rv, int pitch, int pitchc, size_t stride_y, size_t stride_c, int mbw, int mbh,
                       int qp, int alpha_off, int beta_off, const b2_mbinfo_t *__restrict__ info, int *__restrict__ gflags)
{
    __shared__ K8Warp s_warp[WARPS];
    __shared__ int s_flag[WARPS];         // s_flag[w]: macroblocks finished, over all its rows so far, by the warp that produces for warp w
    const int frame = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncta = (int)gridDim.x, crank = (int)blockIdx.x;
    const int gwarp = crank * WARPS + warp, nwarps = ncta * WARPS;
    // gflag[c]: progress of the LAST warp of CTA c-1 (the producer of CTA c's first warp; CTA 0's is the last CTA's, for frames
    // with more rows than warps).  CTAs of one frame are consecutive in dispatch order, so a producer CTA is resident no later
    // than its consumer.
    int *const gflag = gflags + (size_t)frame * 8;
    const b2_mbinfo_t *finfo = info + (size_t)frame * mbw * mbh;
    uint8_t *const py = ry + frame * stride_y, *const pu = ru + frame * stride_c, *const pv = rv + frame * stride_c;
    const int qpc = chroma_qp(qp);
    FiltConst fy, fcc;
    const int ia = clip3(0, 51, qp + 2 * alpha_off), ib = clip3(0, 51, qp + 2 * beta_off);
    const int iac = clip3(0, 51, qpc + 2 * alpha_off), ibc = clip3(0, 51, qpc + 2 * beta_off);
    fy.alpha = c_alpha[ia]; fy.beta = c_beta[ib]; fcc.alpha = c_alpha[iac]; fcc.beta = c_beta[ibc];
#pragma unroll
    for (int i = 0; i < 3; i++) { fy.tc0[i] = c_tc0[ia][i]; fcc.tc0[i] = c_tc0[iac][i]; }
    if (threadIdx.x < WARPS) s_flag[threadIdx.x] = 0;
    __syncthreads();

    K8Warp &ws = s_warp[warp];
    // the consumer of this warp's progress is the warp that owns the next row: the next warp of this CTA (shared-memory flag)
    // or, for the last warp, the first warp of the next CTA (global flag); likewise on the consuming side
    const bool prod_remote = warp == 0, cons_remote = warp == WARPS - 1;
    int *const cons_gflag = &gflag[(crank + 1) % ncta];
    const int *const prod_gflag = &gflag[crank];
    // this lane's boundary-strength job: lanes 0-15 vertical edges, 16-31 horizontal edges; lane -> (edge e, segment k)
    const int dir = lane >> 4, e = (lane >> 2) & 3, k = lane & 3;
    // own-sample words: lane -> luma words (row l>>1, words 2*(l&1), 2*(l&1)+1), chroma word (plane l>>4, row (l>>1)&7, word l&1)
    const int oy_row = lane >> 1, oy_w = (lane & 1) * 2;
    const int oc_p = lane >> 4, oc_row = (lane >> 1) & 7, oc_w = lane & 1;

    int round = 0;
    for (int mby = gwarp; mby < mbh; mby += nwarps, round++) {
        const int prod_base = ((mby - 1) / nwarps) * mbw;      // what the producer had published before it started row mby-1
        const int my_base = round * mbw;
        uint8_t *const rowy = py + (size_t)(B2_PAD + mby * 16) * pitch + B2_PAD;
        uint8_t *const rowu = pu + (size_t)(B2_PADC + mby * 8) * pitchc + B2_PADC;
        uint8_t *const rowv = pv + (size_t)(B2_PADC + mby * 8) * pitchc + B2_PADC;
        // ---- prologue: own samples and boundary strengths of macroblock 0 ----
        uint32_t n0, n1, n2;
        MbLite lq, lp;
        {
            const uint32_t *g = (const uint32_t *)(rowy + (size_t)oy_row * pitch) + oy_w;
            n0 = g[0]; n1 = g[1];
            n2 = *((const uint32_t *)((oc_p ? rowv : rowu) + (size_t)oc_row * pitchc) + oc_w);
            const b2_mbinfo_t *mq = &finfo[mby * mbw];
            lq = load_lite(mq);
            lp = (e == 0 && dir == 1 && mby > 0) ? load_lite(mq - mbw) : lq;
        }
        for (int mbx = 0; mbx < mbw; mbx++) {
            // ---- own samples (prefetched) -> tile; boundary strengths from the prefetched records ----
            *(uint32_t *)&ws.y[(oy_row + 4) * LP + 4 + oy_w * 4] = n0;
            *(uint32_t *)&ws.y[(oy_row + 4) * LP + 8 + oy_w * 4] = n1;
            *(uint32_t *)&ws.c[oc_p][(oc_row + 2) * CP + 4 + oc_w * 4] = n2;
            {
                const bool edge_ok = !(e == 0 && (dir == 0 ? mbx == 0 : mby == 0)) && !((e & 1) && ((lq.pt >> 8) & 255u));
                int bs = 0;
                if (edge_ok) {
                    const int pbx = dir == 0 ? (e == 0 ? 3 : e - 1) : k, pby = dir == 0 ? k : (e == 0 ? 3 : e - 1);
                    const int qbx = dir == 0 ? e : k, qby = dir == 0 ? k : e;
                    bs = lite_bs(lp, pbx, pby, lq, qbx, qby, e == 0);
                }
                ws.bs[dir][e][k] = (int8_t)bs;
            }
            // ---- prefetch the next macroblock: own samples + decision records (nothing in this kernel writes them before) ----
            if (mbx + 1 < mbw) {
                const uint32_t *g = (const uint32_t *)(rowy + (size_t)oy_row * pitch + (mbx + 1) * 16) + oy_w;
                n0 = g[0]; n1 = g[1];
                n2 = *((const uint32_t *)((oc_p ? rowv : rowu) + (size_t)oc_row * pitchc + (mbx + 1) * 8) + oc_w);
                const b2_mbinfo_t *mq = &finfo[mby * mbw + mbx + 1];
                const MbLite cur = lq;                           // becomes the left neighbour
                lq = load_lite(mq);
                lp = e != 0 ? lq : (dir == 0 ? cur : (mby > 0 ? load_lite(mq - mbw) : lq));
            }
            // ---- the four rows above: final once row mby-1 has finished macroblock mbx+1 ----
            if (mby > 0) {
                const int need = prod_base + min(mbx + 2, mbw);
                if (prod_remote) { while (ld_flag_global(prod_gflag) < need) __nanosleep(100); }
                else { while (*(volatile int *)&s_flag[warp] < need) __nanosleep(40); }     // sleeping keeps the issue slots for co-resident kernels
                __threadfence();
                if (lane < 16) {
                    const int r = lane >> 2, w = lane & 3;        // rows -4..-1, words 0..3
                    *(uint32_t *)&ws.y[r * LP + 4 + w * 4] = __ldcg((const uint32_t *)(rowy + (ptrdiff_t)(r - 4) * pitch + mbx * 16) + w);
                } else if (lane < 24) {
                    const int l = lane - 16, p = l >> 2, r = (l >> 1) & 1, w = l & 1;     // rows -2..-1, words 0..1
                    *(uint32_t *)&ws.c[p][r * CP + 4 + w * 4] = __ldcg((const uint32_t *)((p ? rowv : rowu) + (ptrdiff_t)(r - 2) * pitchc + mbx * 8) + w);
                }
            }
            __syncwarp();
            // ---- vertical edges: luma lane = row, chroma lane-16 = (plane, row) ----
            if (lane < 16) {
                uint8_t *row = &ws.y[(lane + 4) * LP + 4];
#pragma unroll
                for (int ee = 0; ee < 4; ee++) filter_line(row + 4 * ee, 1, ws.bs[0][ee][lane >> 2], fy, false);
            } else {
                const int l = lane - 16, p = l >> 3, r = l & 7;
                uint8_t *row = &ws.c[p][(r + 2) * CP + 4];
#pragma unroll
                for (int ee = 0; ee < 2; ee++) filter_line(row + 4 * ee, 1, ws.bs[0][2 * ee][r >> 1], fcc, true);
            }
            __syncwarp();
            // ---- horizontal edges: luma lane = column, chroma lane-16 = (plane, column) ----
            if (lane < 16) {
                uint8_t *col = &ws.y[4 * LP + lane + 4];
#pragma unroll
                for (int ee = 0; ee < 4; ee++) filter_line(col + 4 * ee * LP, LP, ws.bs[1][ee][lane >> 2], fy, false);
            } else {
                const int l = lane - 16, p = l >> 3, x = l & 7;
                uint8_t *col = &ws.c[p][2 * CP + x + 4];
#pragma unroll
                for (int ee = 0; ee < 2; ee++) filter_line(col + 4 * ee * CP, CP, ws.bs[1][2 * ee][x >> 1], fcc, true);
            }
            __syncwarp();
            // ---- write back what can have changed: own MB, 3 (luma) / 1 (chroma) lines into the top MB, one word into the left MB ----
            uint8_t *gy = rowy + mbx * 16;
            for (int i = lane; i < 95; i += 32) {                  // luma rows -3..15 (19 rows) x 5 words
                const int r = i / 5 + 1, w = i - (i / 5) * 5;       // tile row r = y + 4, y = -3..15
                const int y = r - 4;
                if (w == 0 && (mbx == 0 || y < 0)) continue;        // nothing left of the picture; corner block is never modified
                if (y < 0 && mby == 0) continue;
                *(uint32_t *)(gy + (ptrdiff_t)y * pitch + (w - 1) * 4) = *(const uint32_t *)&ws.y[r * LP + w * 4];
            }
            for (int i = lane; i < 54; i += 32) {                  // chroma rows -1..7 (9 rows) x 3 words x 2 planes
                const int p = i / 27, kk = i - p * 27, r = kk / 3 + 1, w = kk - (kk / 3) * 3;
                const int y = r - 2;
                if (w == 0 && (mbx == 0 || y < 0)) continue;
                if (y < 0 && mby == 0) continue;
                uint8_t *gc = (p ? rowv : rowu) + mbx * 8;
                *(uint32_t *)(gc + (ptrdiff_t)y * pitchc + (w - 1) * 4) = *(const uint32_t *)&ws.c[p][r * CP + w * 4];
            }
            // ---- publish: every lane's stores are visible GPU-wide before the flag moves ----
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                if (cons_remote) st_flag_global(cons_gflag, my_base + mbx + 1);
                else *(volatile int *)&s_flag[warp + 1] = my_base + mbx + 1;
            }
            // ---- carry the four rightmost columns over as the next macroblock's left neighbour ----
            if (lane < 16) *(uint32_t *)&ws.y[(lane + 4) * LP] = *(const uint32_t *)&ws.y[(lane + 4) * LP + 16];
            else *(uint32_t *)&ws.c[(lane - 16) >> 3][(((lane - 16) & 7) + 2) * CP] = *(const uint32_t *)&ws.c[(lane - 16) >> 3][(((lane - 16) & 7) + 2) * CP + 8];
            __syncwarp();
        }
    }
}

}  // namespace

int b2_launch_deblock(uint8_t *const rec[3], int pitch, int pitchc, size_t stride_y, size_t stride_c, int mbw, int mbh,
                      int nframes, int qp, int alpha_off, int beta_off, const b2_mbinfo_t *d_info, int *d_flags, cudaStream_t st)
{
    static const bool wavefront = getenv("B2_K8_WAVEFRONT") && atoi(getenv("B2_K8_WAVEFRONT")) != 0;      // A/B measurements only
    int ncta = 1;
    if (wavefront) {
        const int maxdiag = mbh < (mbw + 1) / 2 ? mbh : (mbw + 1) / 2;
        while (ncta < 8 && ncta * K8_WARPS < maxdiag) ncta *= 2;
    } else {
        ncta = (mbh + K8_WARPS - 1) / K8_WARPS;                   // one warp per macroblock row; beyond 128 rows warps take several
        if (ncta > 8) ncta = 8;
    }
    if (!wavefront) {
        // plain launch: no cluster to place, so a frame's few CTAs slot in beside whatever else runs on the GPU
        B2_CUDA_OK(cudaMemsetAsync(d_flags, 0, (size_t)nframes * 8 * sizeof(int), st));
        k8_deblock_rows_kernel<K8_WARPS><<<dim3(ncta, nframes), K8_WARPS * 32, 0, st>>>(rec[0], rec[1], rec[2], pitch, pitchc, stride_y, stride_c,
                                                                                        mbw, mbh, qp, alpha_off, beta_off, d_info, d_flags);
        B2_CUDA_OK(cudaGetLastError());
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncta, nframes, 1);
    cfg.blockDim = dim3(K8_WARPS * 32, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ncta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    B2_CUDA_OK(cudaLaunchKernelEx(&cfg, k8_deblock_kernel, rec[0], rec[1], rec[2], pitch, pitchc, stride_y, stride_c, mbw, mbh, qp,
                                  alpha_off, beta_off, d_info));
    return 0;
}
