// k2_me_subpel.cu -- K2: half- then quarter-pel refinement of the full-pel motion vector by SATD.
//
// Replaces the sub-pel half of x264's motion estimation (behind x264_encoder_encode,
// av_encode.c:970); bit-exact against oracle/b2o_me.c:b2o_me_subpel (9 half-pel candidates,
// centre first, then the 8 quarter-pel neighbours of the winner; cost = SATD16x16 +
// lambda*mvbits(mv - pmv); first minimum in scan order wins).
//
// One CTA per macroblock.  The 22x22 reference patch around the full-pel winner is staged in shared
// memory once; the three half-sample planes b (horizontal), h (vertical) and j (centre, from the
// UNROUNDED horizontal intermediates, as 8.4.2.2.1 requires) are built once and every one of the
// 17 candidates is then a two-plane average -- the 6-tap filter runs ~1,300 times per MB instead
// of ~50,000 times if each candidate were interpolated on its own.
// Bound: integer ALU; algorithmic work = 17 candidates x 16 SATD4x4 per MB (~10x below K1).
#include "b2_h264.cuh"

namespace {

constexpr int K2_THREADS = 160;

struct K2Smem {
    uint8_t G[22][24];        // integer samples, rows/cols -3..18 around the full-pel position
    int16_t B1[22][18];       // unrounded horizontal half samples: rows -3..18, cols -1..15
    uint8_t Bq[18][20];       // b: rows -1..16, cols -1..15
    uint8_t H[17][20];        // h: rows -1..15, cols -1..16
    uint8_t J[17][20];        // j: rows -1..15, cols -1..15
    uint8_t cur[16][16];
    uint32_t cost[9];
    int best;
};

// sample at integer base (X,Y) in [-1,16] with quarter-pel fraction (fx,fy), from the planes
__device__ __forceinline__ int plane_sample(const K2Smem &s, int X, int Y, int fx, int fy)
{
#define PG(x, y) ((int)s.G[(y) + 3][(x) + 3])
#define PB(x, y) ((int)s.Bq[(y) + 1][(x) + 1])
#define PH(x, y) ((int)s.H[(y) + 1][(x) + 1])
#define PJ(x, y) ((int)s.J[(y) + 1][(x) + 1])
#define AV(a, b) (((a) + (b) + 1) >> 1)
    switch (fy * 4 + fx) {
    case 0: return PG(X, Y);
    case 1: return AV(PG(X, Y), PB(X, Y));
    case 2: return PB(X, Y);
    case 3: return AV(PG(X + 1, Y), PB(X, Y));
    case 4: return AV(PG(X, Y), PH(X, Y));
    case 5: return AV(PB(X, Y), PH(X, Y));
    case 6: return AV(PB(X, Y), PJ(X, Y));
    case 7: return AV(PB(X, Y), PH(X + 1, Y));
    case 8: return PH(X, Y);
    case 9: return AV(PH(X, Y), PJ(X, Y));
    case 10: return PJ(X, Y);
    case 11: return AV(PJ(X, Y), PH(X + 1, Y));
    case 12: return AV(PG(X, Y + 1), PH(X, Y));
    case 13: return AV(PH(X, Y), PB(X, Y + 1));
    case 14: return AV(PJ(X, Y), PB(X, Y + 1));
    default: return AV(PH(X + 1, Y), PB(X, Y + 1));
    }
#undef PG
#undef PB
#undef PH
#undef PJ
#undef AV
}

__device__ __constant__ int8_t c_subpel_off[9][2] = {{0, 0}, {-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};

// SATD of 4x4 block `blk` (raster 0..15 inside the MB) for the candidate displaced by (cx,cy)
// quarter-pels from the full-pel position
__device__ __forceinline__ uint32_t cand_block_satd(const K2Smem &s, int blk, int cx, int cy)
{
    const int ix = cx >> 2, iy = cy >> 2, fx = cx & 3, fy = cy & 3;
    const int bx = (blk & 3) * 4, by = (blk >> 2) * 4;
    int d[16];
#pragma unroll
    for (int y = 0; y < 4; y++)
#pragma unroll
        for (int x = 0; x < 4; x++)
            d[y * 4 + x] = (int)s.cur[by + y][bx + x] - plane_sample(s, bx + x + ix, by + y + iy, fx, fy);
    return b2::satd4x4(d);
}

__global__ void __launch_bounds__(K2_THREADS)
k2_me_subpel_kernel(const uint8_t *__restrict__ cur, const uint8_t *__restrict__ ref, int pitch, size_t plane_stride,
                    int mbw, int mbh, const b2_mv_t *__restrict__ mv_full, const b2_mv_t *__restrict__ pmv,
                    int lambda, int subpel, b2_mv_t *__restrict__ mv_out, uint32_t *__restrict__ cost_out)
{
    __shared__ K2Smem s;
    const int tid = threadIdx.x;
    const int mbx = blockIdx.x, mby = blockIdx.y, frame = blockIdx.z;
    const size_t mbi = ((size_t)frame * mbh + mby) * mbw + mbx;
    const b2_mv_t mvf = mv_full[mbi];
    b2_mv_t pm = {0, 0};
    if (pmv) pm = pmv[mbi];

    const uint8_t *cplane = cur + frame * plane_stride + (size_t)(B2_PAD + mby * 16) * pitch + B2_PAD + mbx * 16;
    const uint8_t *rplane = ref + frame * plane_stride + (size_t)(B2_PAD + mby * 16 + mvf.y - 3) * pitch + B2_PAD +
                            mbx * 16 + mvf.x - 3;
    for (int i = tid; i < 22 * 22; i += K2_THREADS) {
        int r = i / 22, c = i - r * 22;
        s.G[r][c] = rplane[(size_t)r * pitch + c];
    }
    for (int i = tid; i < 64; i += K2_THREADS) {
        int r = i >> 2, c = i & 3;
        *(uint32_t *)&s.cur[r][c * 4] = *(const uint32_t *)(cplane + (size_t)r * pitch + c * 4);
    }
    if (tid < 9) s.cost[tid] = 0;
    __syncthreads();

    // horizontal unrounded half samples (all 22 rows) and vertical half samples
    for (int i = tid; i < 22 * 17 + 17 * 18; i += K2_THREADS) {
        if (i < 22 * 17) {
            int r = i / 17, c = i - r * 17;               // row r <-> Y=r-3, col c <-> X=c-1 ; G col = X+3 = c+2
            const uint8_t *q = &s.G[r][c + 2];
            s.B1[r][c] = (int16_t)b2::tap6(q[-2], q[-1], q[0], q[1], q[2], q[3]);
        } else {
            int k = i - 22 * 17;
            int r = k / 18, c = k - r * 18;               // Y=r-1, X=c-1 ; G[Y+3][X+3] = G[r+2][c+2]
            int v = b2::tap6(s.G[r][c + 2], s.G[r + 1][c + 2], s.G[r + 2][c + 2], s.G[r + 3][c + 2], s.G[r + 4][c + 2],
                             s.G[r + 5][c + 2]);
            s.H[r][c] = (uint8_t)b2_clip255((v + 16) >> 5);
        }
    }
    __syncthreads();
    for (int i = tid; i < 18 * 17 + 17 * 17; i += K2_THREADS) {
        if (i < 18 * 17) {
            int r = i / 17, c = i - r * 17;               // Y=r-1 -> B1 row Y+3 = r+2
            s.Bq[r][c] = (uint8_t)b2_clip255((s.B1[r + 2][c] + 16) >> 5);
        } else {
            int k = i - 18 * 17;
            int r = k / 17, c = k - r * 17;               // Y=r-1 -> B1 rows Y-2+3 .. Y+3+3 = r .. r+5
            int v = b2::tap6(s.B1[r][c], s.B1[r + 1][c], s.B1[r + 2][c], s.B1[r + 3][c], s.B1[r + 4][c], s.B1[r + 5][c]);
            s.J[r][c] = (uint8_t)b2_clip255((v + 512) >> 10);
        }
    }
    __syncthreads();

    // stage 1: centre + 8 half-pel neighbours (offsets x2 quarter units); without sub-pel: centre only
    const int n1 = subpel ? 9 : 1;
    if (tid < n1 * 16) {
        int cand = tid >> 4, blk = tid & 15;
        uint32_t v = cand_block_satd(s, blk, 2 * c_subpel_off[cand][0], 2 * c_subpel_off[cand][1]);
        atomicAdd(&s.cost[cand], v);
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t best = 0xffffffffu; int bi = 0;
        for (int k = 0; k < n1; k++) {
            int mx = mvf.x * 4 + 2 * c_subpel_off[k][0], my = mvf.y * 4 + 2 * c_subpel_off[k][1];
            uint32_t c = s.cost[k] + (uint32_t)(lambda * (b2_mvbits(mx - pm.x) + b2_mvbits(my - pm.y)));
            if (c < best) { best = c; bi = k; }
        }
        s.best = bi;
        s.cost[0] = best;
        for (int k = 1; k < 9; k++) s.cost[k] = 0;
    }
    __syncthreads();
    const int hx = 2 * c_subpel_off[s.best][0], hy = 2 * c_subpel_off[s.best][1];
    if (subpel) {
        if (tid < 8 * 16) {
            int cand = 1 + (tid >> 4), blk = tid & 15;
            uint32_t v = cand_block_satd(s, blk, hx + c_subpel_off[cand][0], hy + c_subpel_off[cand][1]);
            atomicAdd(&s.cost[cand], v);
        }
        __syncthreads();
    }
    if (tid == 0) {
        uint32_t best = s.cost[0];
        int bx = mvf.x * 4 + hx, by = mvf.y * 4 + hy;
        if (subpel)
            for (int k = 1; k < 9; k++) {
                int mx = mvf.x * 4 + hx + c_subpel_off[k][0], my = mvf.y * 4 + hy + c_subpel_off[k][1];
                uint32_t c = s.cost[k] + (uint32_t)(lambda * (b2_mvbits(mx - pm.x) + b2_mvbits(my - pm.y)));
                if (c < best) { best = c; bx = mx; by = my; }
            }
        b2_mv_t o;
        o.x = (int16_t)bx; o.y = (int16_t)by;
        mv_out[mbi] = o;
        cost_out[mbi] = best;
    }
}

}  // namespace

int b2_launch_me_subpel(const uint8_t *d_cur, const uint8_t *d_ref, int pitch, size_t plane_stride, int mbw, int mbh,
                        int nframes, const b2_mv_t *d_mv_full, const b2_mv_t *d_pmv, int lambda, int subpel,
                        b2_mv_t *d_mv_out, uint32_t *d_cost_out, cudaStream_t st)
{
    dim3 grid(mbw, mbh, nframes);
    k2_me_subpel_kernel<<<grid, K2_THREADS, 0, st>>>(d_cur, d_ref, pitch, plane_stride, mbw, mbh, d_mv_full, d_pmv,
                                                     lambda, subpel, d_mv_out, d_cost_out);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}
