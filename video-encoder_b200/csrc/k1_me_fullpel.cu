// k1_me_fullpel.cu -- K1: exhaustive full-pel SAD motion search, one MV per 16x16 macroblock.
//
// Replaces the full-pel half of the motion estimation that the reference reaches through
// x264_encoder_encode() (av_encode.c:970); bit-exact against oracle/b2o_me.c:b2o_me_fullpel.
//
// Bound: integer ALU (VABSDIFF4.U8.ACC issue rate).  Algorithmic work per macroblock:
// (2R+1)^2 candidates x 256 pixel-SADs = (2R+1)^2 x 64 VABSDIFF4 lane-instructions.
//
// Design:
//  * One CTA searches a strip of NMB = 6 horizontally adjacent macroblocks of one frame.  The
//    (16 NMB + 2R) x (16+2R) search window of the *padded* reference plane and the 16 NMB x 16 current
//    tile are fetched by TMA (cp.async.bulk.tensor.3d), never leaving the allocation because
//    every plane carries a 64-px replicated border.
//  * Operand supply, not the ALU, is the first limiter (32 shared-memory words/clk/SM vs 64
//    VABSDIFF4 lanes/clk/SM).  Two reuse tricks remove it:
//      - the window is expanded once into "one 32-bit word per pixel position"
//        (word x = pixels x..x+3), so every byte alignment is a plain conflict-free LDS.32;
//      - register tiling: a lane owns one dx and K consecutive dy; the current macroblock sits
//        in 64 registers, each reference row is loaded once (4 LDS.32) and applied to up to K
//        current rows (K accumulators) -> (16+K-1)*4 loads per 64*K VABSDIFF4.
//  * Lanes of a warp are 32 consecutive LANE-TASKS (mb, dy-group, dx).  Inside one (mb, dy-group) segment that is 32 consecutive
//    dx = 32 consecutive words: one wavefront.  A segment is 2R+1 = 65 (33) tasks long, not a multiple of 32, so about every other
//    warp-task straddles two segments whose rows start at the same bank (EXP_PITCH % 32 == 0): its sweep LDS.32 then take two
//    wavefronts.  ncu, per instruction (profiles/r2_k1_lds_bank_conflicts.txt): every sweep LDS averages 1.475 wavefronts, together
//    17.8 M of the kernel's 18.4 M excess wavefronts per 4-frame launch; the remaining 0.6 M are the ATOMS.MIN on the best keys.
//    The shared-memory pipe still runs at only 0.46 wavefronts/clk/SM while the ALU pipe is 92 % busy, so this is not the limiter --
//    and the fix was measured to cost more than it saves: a pitch with K * pitch = 1 (mod 32) lines the two segments' banks up, but
//    such a pitch is odd, so the expansion can no longer store 16 bytes per lane; bit-identical, 0.889 -> 0.868 of the VABSDIFF4 peak
//    with word-wise expansion alone and 0.853 with the padded pitch (+-16: 0.836 -> 0.787; scripts/k1_exppad_probe.py,
//    profiles/r2_k1_exppad_ab.txt).  The conflicts stay.
//  * Winner: key = (cost << 13) | scan_index, CREDUX.MIN over the warp, atomicMin in smem.
//    Lowest scan index wins ties by construction, exactly like the oracle's strict '<'.
#include <stdlib.h>
#include <atomic>
#include "b2_common.cuh"

namespace {

// Strip width and residency, tuned on the B200 (scripts/k1_variants.sh + k1_variant_probe.py, 1080p +-32, 32 frames):
//   NMB x CTAs/SM :  8x2 0.859   10x2 0.874   12x2 0.879   5x3 0.885   6x3 0.888   7x3 0.881   4x3 0.871   of the VABSDIFF4 peak.
// Three resident CTAs (80 registers, 67 KB of shared memory each) keep the ALU pipe fed while one of them waits for its
// TMA window / expands it; a wider strip only amortises the halo.
#ifndef B2_K1_NMB
#define B2_K1_NMB 6
#endif
#ifndef B2_K1_MINCTAS
#define B2_K1_MINCTAS 3
#endif
#ifndef B2_K1_SMEM_PAD
#define B2_K1_SMEM_PAD 0    // extra dynamic shared memory per CTA (bytes): lowers K1's residency without touching its register budget,
#endif                      // so that other stream groups' kernels can co-reside (experiment knob, scripts/k1_variants.sh)
#ifndef B2_K1_NMB16
#define B2_K1_NMB16 10      // +-16: strips of 10 MBs.  Swept on the B200 with 192/256/320/384 threads (scripts/k1_r16_variant_probe.py,
#endif                      // profiles/r2_k1_r16_strip_sweep.txt): 8 -> 0.821, 10 -> 0.838, 12 -> 0.817, 14 -> 0.829 of the VABSDIFF4 peak at 720p

template <int R> struct K1Cfg;
// NMB = macroblocks per CTA strip (tuning: scripts/k1_variants.sh)
template <> struct K1Cfg<32> { static constexpr int K = 13, NG = 5, NMB = B2_K1_NMB; };     // 65 = 5 x 13
template <> struct K1Cfg<16> { static constexpr int K = 11, NG = 3, NMB = B2_K1_NMB16; };   // 33 = 3 x 11

template <int R> struct K1Smem {
    static constexpr int NMB = K1Cfg<R>::NMB;
    static constexpr int ND = 2 * R + 1;
    static constexpr int WIN_W = NMB * 16 + 2 * R;      // bytes per raw window row
    static constexpr int WIN_H = 16 + 2 * R;
    static constexpr int RAW_BYTES = WIN_W * WIN_H;
    static constexpr int EXP_PITCH = WIN_W;             // words per expanded row
    static constexpr int EXP_BYTES = EXP_PITCH * WIN_H * 4;
    static constexpr int CUR_BYTES = NMB * 16 * 16;
    // layout (all 128-B aligned)
    static constexpr int OFF_RAW = 0;
    static constexpr int OFF_CUR = (RAW_BYTES + 16 + 127) & ~127;
    static constexpr int OFF_EXP = OFF_CUR + CUR_BYTES;
    static constexpr int OFF_COSTX = OFF_EXP + EXP_BYTES;
    static constexpr int OFF_COSTY = OFF_COSTX + NMB * ND * 4;
    static constexpr int OFF_BEST = OFF_COSTY + NMB * ND * 4;
    static constexpr int OFF_BAR = (OFF_BEST + NMB * 9 * 4 + 7) & ~7;    // 9 best keys per MB in the partition variant
    static constexpr int TOTAL = OFF_BAR + 8 + 128;     // +128: manual alignment slack
};

// PART (row N1, partitions = 2): the same sweep also yields the SAD of every 8x8 quadrant, hence the best vector of each of
// the nine shape parts (16x16 | 16x8 top,bottom | 8x16 left,right | four 8x8).  The number of VABSDIFF4 is unchanged; a lane
// sweeps the top half of the current MB, packs the two 8x8 SADs of its K candidates into K registers, sweeps the bottom
// half, and then forms 9 sums / keys per candidate (9 running minima, 9 CREDUX.MIN + 9 shared atomicMin per warp task).
template <int R, int NTHREADS, bool PART>
__global__ void __launch_bounds__(NTHREADS, PART ? 2 : B2_K1_MINCTAS)      // the partition variant needs ~90 registers: two CTAs
k1_me_fullpel_kernel(const __grid_constant__ CUtensorMap tm_cur,
                     const __grid_constant__ CUtensorMap tm_ref,
                     int mbw, int mbh, const b2_mv_t *__restrict__ pmv, int lambda,
                     b2_mv_t *__restrict__ mv_out, uint32_t *__restrict__ cost_out,
                     b2_mv_t *__restrict__ mv9_out, uint32_t *__restrict__ cost9_out)
{
    using S = K1Smem<R>;
    constexpr int NWARPS = NTHREADS / 32;
    constexpr int K = K1Cfg<R>::K, NG = K1Cfg<R>::NG, ND = S::ND, NMB = K1Cfg<R>::NMB;
    static_assert(K * NG == ND, "dy groups must tile the search range");

    extern __shared__ uint8_t smem_raw_[];
    uint8_t *smem = smem_raw_ + ((128u - (smem_u32(smem_raw_) & 127u)) & 127u);   // keeps the shared address space
    uint8_t *s_raw = smem + S::OFF_RAW;
    uint8_t *s_cur = smem + S::OFF_CUR;
    uint32_t *s_exp = (uint32_t *)(smem + S::OFF_EXP);
    uint32_t *s_costx = (uint32_t *)(smem + S::OFF_COSTX);
    uint32_t *s_costy = (uint32_t *)(smem + S::OFF_COSTY);
    uint32_t *s_best = (uint32_t *)(smem + S::OFF_BEST);
    uint64_t *s_bar = (uint64_t *)(smem + S::OFF_BAR);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mb0 = blockIdx.x * NMB, mby = blockIdx.y, frame = blockIdx.z;
    const int nmb = min(NMB, mbw - mb0);

    if (tid == 0) {
        mbar_init(s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(s_bar, S::RAW_BYTES + S::CUR_BYTES);
        tma_load_3d(s_raw, &tm_ref, B2_PAD + mb0 * 16 - R, B2_PAD + mby * 16 - R, frame, s_bar);
        tma_load_3d(s_cur, &tm_cur, B2_PAD + mb0 * 16, B2_PAD + mby * 16, frame, s_bar);
    }

    // while the TMA is in flight: motion-vector cost tables and best-key init
    const size_t mb_base = ((size_t)frame * mbh + mby) * mbw + mb0;
    for (int i = tid; i < NMB * ND; i += NTHREADS) {
        int m = i / ND, d = i - m * ND;
        int px = 0, py = 0;
        if (pmv != nullptr && m < nmb) { b2_mv_t p = pmv[mb_base + m]; px = p.x; py = p.y; }
        // pre-shifted so that key = (SAD << 13) + kx + ky = ((SAD + costx + costy) << 13) | scan index
        s_costx[i] = ((uint32_t)(lambda * b2_mvbits(4 * (d - R) - px)) << 13) + (uint32_t)d;
        s_costy[i] = ((uint32_t)(lambda * b2_mvbits(4 * (d - R) - py)) << 13) + (uint32_t)(d * ND);
    }
    if (tid < NMB * (PART ? 9 : 1)) s_best[tid] = 0xffffffffu;

    mbar_wait(s_bar, 0);

    // expand: word x of a row = pixels x..x+3 (little endian), all 4 byte alignments
    {
        constexpr int WPR = S::WIN_W / 4;             // aligned words per raw row
        const uint32_t *raw32 = (const uint32_t *)s_raw;
        for (int i = tid; i < WPR * S::WIN_H; i += NTHREADS) {
            int r = i / WPR, j = i - r * WPR;
            uint32_t lo = raw32[i], hi = raw32[i + 1];    // last word of the buffer: 16 B slack
            uint4 o;
            o.x = lo;
            o.y = __funnelshift_r(lo, hi, 8);
            o.z = __funnelshift_r(lo, hi, 16);
            o.w = __funnelshift_r(lo, hi, 24);
            *(uint4 *)(s_exp + r * S::EXP_PITCH + 4 * j) = o;
        }
    }
    __syncthreads();

    const int total = nmb * NG * ND;                  // lane-tasks: (mb, dy-group, dx)
    for (int t0 = warp * 32; t0 < total; t0 += NWARPS * 32) {
        const int t = t0 + lane;
        const bool active = t < total;
        const int tt = active ? t : t0;               // idle lanes shadow lane 0's task
        const int m = tt / (NG * ND);
        const int rem = tt - m * (NG * ND);
        const int g = rem / ND;
        const int dxi = rem - g * ND;

        if constexpr (PART) {
            uint32_t top[K];                               // SAD(top-left 8x8) | SAD(top-right 8x8) << 16
            uint32_t best[9];
#pragma unroll
            for (int p = 0; p < 9; p++) best[p] = 0xffffffffu;
            const uint32_t kx = s_costx[m * ND + dxi];
            const uint32_t *ky = s_costy + m * ND + g * K;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                uint32_t cur[32];                          // 8 rows of the current macroblock
                {
                    const uint4 *c4 = (const uint4 *)(s_cur + m * 16) + half * 8 * (NMB * 16 / 16);
#pragma unroll
                    for (int y = 0; y < 8; y++) {
                        uint4 v = c4[y * (NMB * 16 / 16)];
                        cur[y * 4 + 0] = v.x; cur[y * 4 + 1] = v.y; cur[y * 4 + 2] = v.z; cur[y * 4 + 3] = v.w;
                    }
                }
                uint32_t accl[K], accr[K];
#pragma unroll
                for (int k = 0; k < K; k++) accl[k] = accr[k] = 0;
                const uint32_t *wp = s_exp + (g * K + half * 8) * S::EXP_PITCH + m * 16 + dxi;
#pragma unroll
                for (int r = 0; r < K + 7; r++) {
                    const uint32_t w0 = wp[r * S::EXP_PITCH + 0];
                    const uint32_t w1 = wp[r * S::EXP_PITCH + 4];
                    const uint32_t w2 = wp[r * S::EXP_PITCH + 8];
                    const uint32_t w3 = wp[r * S::EXP_PITCH + 12];
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const int y = r - k;
                        if (y >= 0 && y < 8) {
                            accl[k] = vsad4_acc(w0, cur[y * 4 + 0], accl[k]);
                            accl[k] = vsad4_acc(w1, cur[y * 4 + 1], accl[k]);
                            accr[k] = vsad4_acc(w2, cur[y * 4 + 2], accr[k]);
                            accr[k] = vsad4_acc(w3, cur[y * 4 + 3], accr[k]);
                        }
                    }
                }
                if (half == 0) {
#pragma unroll
                    for (int k = 0; k < K; k++) top[k] = accl[k] | (accr[k] << 16);
                } else if (active) {
                    uint32_t pend[9];                          // keys of the previous candidate: merged two at a time (VIMNMX3)
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const uint32_t tl = top[k] & 0xffffu, tr = top[k] >> 16, bl = accl[k], br = accr[k];
                        const uint32_t kc = kx + ky[k];
                        const uint32_t st = tl + tr, sb = bl + br;
                        const uint32_t c[9] = {(st + sb) * 8192u + kc, st * 8192u + kc, sb * 8192u + kc, (tl + bl) * 8192u + kc,
                                               (tr + br) * 8192u + kc, tl * 8192u + kc,  tr * 8192u + kc, bl * 8192u + kc, br * 8192u + kc};
#pragma unroll
                        for (int p = 0; p < 9; p++) {
                            if (k & 1) best[p] = __vimin3_u32(best[p], pend[p], c[p]);
                            else if (k == K - 1) best[p] = min(best[p], c[p]);
                            else pend[p] = c[p];
                        }
                    }
                }
            }
            const int m0 = __shfl_sync(0xffffffffu, m, 0);
            if (__all_sync(0xffffffffu, m == m0)) {
#pragma unroll
                for (int p = 0; p < 9; p++) {
                    const uint32_t wmin = __reduce_min_sync(0xffffffffu, best[p]);
                    if (lane == 0) atomicMin(&s_best[m0 * 9 + p], wmin);
                }
            } else if (active) {
#pragma unroll
                for (int p = 0; p < 9; p++) atomicMin(&s_best[m * 9 + p], best[p]);
            }
        } else {
        // current macroblock -> 64 registers
        uint32_t cur[64];
        {
            const uint4 *c4 = (const uint4 *)(s_cur + m * 16);
#pragma unroll
            for (int y = 0; y < 16; y++) {
                uint4 v = c4[y * (NMB * 16 / 16)];
                cur[y * 4 + 0] = v.x; cur[y * 4 + 1] = v.y; cur[y * 4 + 2] = v.z; cur[y * 4 + 3] = v.w;
            }
        }
        uint32_t acc[K];
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = 0;

        const uint32_t *wp = s_exp + (g * K) * S::EXP_PITCH + m * 16 + dxi;
#pragma unroll
        for (int r = 0; r < K + 15; r++) {
            const uint32_t w0 = wp[r * S::EXP_PITCH + 0];
            const uint32_t w1 = wp[r * S::EXP_PITCH + 4];
            const uint32_t w2 = wp[r * S::EXP_PITCH + 8];
            const uint32_t w3 = wp[r * S::EXP_PITCH + 12];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int y = r - k;                  // current row that meets reference row r at dy = g*K+k
                if (y >= 0 && y < 16) {
                    acc[k] = vsad4_acc(w0, cur[y * 4 + 0], acc[k]);
                    acc[k] = vsad4_acc(w1, cur[y * 4 + 1], acc[k]);
                    acc[k] = vsad4_acc(w2, cur[y * 4 + 2], acc[k]);
                    acc[k] = vsad4_acc(w3, cur[y * 4 + 3], acc[k]);
                }
            }
        }

        uint32_t key = 0xffffffffu;
        if (active) {
            const uint32_t kx = s_costx[m * ND + dxi];
            const uint32_t *ky = s_costy + m * ND + g * K;
            // three-input minimum (VIMNMX3): half as many ALU-pipe instructions beside the VABSDIFF4 stream
#pragma unroll
            for (int k = 0; k + 1 < K; k += 2) key = __vimin3_u32(key, acc[k] * 8192u + (kx + ky[k]), acc[k + 1] * 8192u + (kx + ky[k + 1]));
            if (K & 1) key = min(key, acc[K - 1] * 8192u + (kx + ky[K - 1]));
        }
        const int m0 = __shfl_sync(0xffffffffu, m, 0);
        if (__all_sync(0xffffffffu, m == m0)) {
            const uint32_t wmin = __reduce_min_sync(0xffffffffu, key);
            if (lane == 0) atomicMin(&s_best[m0], wmin);
        } else if (active) {
            atomicMin(&s_best[m], key);
        }
        }
    }
    __syncthreads();

    if (tid < nmb * (PART ? 9 : 1)) {
        const uint32_t key = s_best[tid];
        const int idx = (int)(key & 8191u);
        const int dyi = idx / ND, dxi = idx - dyi * ND;
        b2_mv_t mv;
        mv.x = (int16_t)(dxi - R);
        mv.y = (int16_t)(dyi - R);
        if constexpr (PART) {
            const int mm = tid / 9, p = tid - mm * 9;
            mv9_out[(mb_base + mm) * 9 + p] = mv;
            cost9_out[(mb_base + mm) * 9 + p] = key >> 13;
            if (p == 0) { mv_out[mb_base + mm] = mv; cost_out[mb_base + mm] = key >> 13; }
        } else {
            mv_out[mb_base + tid] = mv;
            cost_out[mb_base + tid] = key >> 13;
        }
    }
}

// ---- persistent, software-pipelined form of the single-vector search ------------------------------------------------------
// One CTA per SM walks a list of strips.  A PRODUCER warp fetches strip k+1 (TMA raw window + current tile), expands the
// window, fills the cost tables and publishes the buffer through an mbarrier while NCW CONSUMER warps still sweep strip k out
// of the other buffer; consumers take warp-tasks from one global sequence (task T -> strip T / TPS), so they drift across the
// strip boundary without a CTA-wide barrier and the ALU pipe never sees a strip prologue.  Each finished task arrives on the
// buffer's "empty" barrier; when all TPS tasks of a strip are in, the producer writes that strip's vectors and refills the
// buffer.  Arithmetic, scan order and tie-break are those of k1_me_fullpel_kernel.
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int R> struct K1PSmem {
    using S1 = K1Smem<R>;
    static constexpr int NMB = S1::NMB, ND = S1::ND;
    static constexpr int TAB = NMB * ND * 4;
    static constexpr int OFF_RAW = 0;
    static constexpr int BUF0 = (S1::RAW_BYTES + 16 + 127) & ~127;
    // per buffer: cur | exp | costx | costy | best
    static constexpr int B_CUR = 0;
    static constexpr int B_EXP = S1::CUR_BYTES;
    static constexpr int B_COSTX = B_EXP + S1::EXP_BYTES;
    static constexpr int B_COSTY = B_COSTX + TAB;
    static constexpr int B_BEST = B_COSTY + TAB;
    static constexpr int BUF_BYTES = (B_BEST + NMB * 4 + 127) & ~127;
    static constexpr int NBUF = 3;                           // strips in flight: one being swept, two prepared ahead
    static constexpr int OFF_BAR = BUF0 + NBUF * BUF_BYTES;  // full[NBUF], empty[NBUF], tma
    static constexpr int TOTAL = OFF_BAR + (2 * NBUF + 1) * 8 + 128;      // +128: manual alignment slack
};

constexpr int K1P_PW = 4;                                   // producer warps
__device__ __forceinline__ void producer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(K1P_PW * 32) : "memory"); }

template <int R, int NCW>
__global__ void __launch_bounds__((NCW + K1P_PW) * 32, 1)
k1_me_fullpel_persistent_kernel(const __grid_constant__ CUtensorMap tm_cur, const __grid_constant__ CUtensorMap tm_ref,
                                int mbw, int mbh, int nframes, const b2_mv_t *__restrict__ pmv, int lambda,
                                b2_mv_t *__restrict__ mv_out, uint32_t *__restrict__ cost_out)
{
    using S = K1Smem<R>;
    using P = K1PSmem<R>;
    constexpr int K = K1Cfg<R>::K, NG = K1Cfg<R>::NG, ND = S::ND, NMB = K1Cfg<R>::NMB;
    constexpr int TPS = (NMB * NG * ND + 31) / 32;              // warp-tasks of a full strip (ragged strips pad with no-ops)

    extern __shared__ uint8_t smem_raw_[];
    uint8_t *smem = smem_raw_ + ((128u - (smem_u32(smem_raw_) & 127u)) & 127u);
    uint8_t *s_raw = smem + P::OFF_RAW;
    constexpr int NBUF = P::NBUF;
    uint64_t *s_full = (uint64_t *)(smem + P::OFF_BAR), *s_empty = s_full + NBUF, *s_tma = s_full + 2 * NBUF;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int spr = (mbw + NMB - 1) / NMB;                       // strips per macroblock row
    const int nstrips = spr * mbh * nframes;
    const int mine = nstrips > (int)blockIdx.x ? (nstrips - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // strips of this CTA

    if (tid == 0) {
        for (int b = 0; b < NBUF; b++) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], TPS); }
        mbar_init(s_tma, 1);
        fence_mbar_init();
    }
    __syncthreads();

    if (warp >= NCW) {
        // ================= producer warps =================
        const int ptid = tid - NCW * 32;                         // 0 .. K1P_PW*32-1
        __shared__ int s_pmv[2][16];                             // predictors of the strip being prepared (NMB <= 16)
        static_assert(NMB <= 16, "strip wider than the predictor staging");
        for (int k = 0; k < mine + NBUF; k++) {
            const int buf = k % NBUF, use = k / NBUF;
            uint8_t *bb = smem + P::BUF0 + buf * P::BUF_BYTES;
            uint32_t *s_best = (uint32_t *)(bb + P::B_BEST);
            if (k >= NBUF) {
                // all tasks of strip k-NBUF have arrived: write its vectors, then the buffer is free
                mbar_wait(&s_empty[buf], (uint32_t)((use - 1) & 1));
                const int sidx = (int)blockIdx.x + (k - NBUF) * (int)gridDim.x;
                const int frame = sidx / (spr * mbh), rem = sidx - frame * (spr * mbh);
                const int mby = rem / spr, mb0 = (rem - mby * spr) * NMB;
                const int nmb = min(NMB, mbw - mb0);
                if (ptid < nmb) {
                    const uint32_t key = s_best[ptid];
                    const int idx = (int)(key & 8191u);
                    const int dyi = idx / ND, dxi = idx - dyi * ND;
                    b2_mv_t mv;
                    mv.x = (int16_t)(dxi - R); mv.y = (int16_t)(dyi - R);
                    const size_t o = ((size_t)frame * mbh + mby) * mbw + mb0 + ptid;
                    mv_out[o] = mv; cost_out[o] = key >> 13;
                }
                producer_sync();                                 // the keys are read before the buffer is refilled
            }
            if (k >= mine) continue;
            const int sidx = (int)blockIdx.x + k * (int)gridDim.x;
            const int frame = sidx / (spr * mbh), rem = sidx - frame * (spr * mbh);
            const int mby = rem / spr, mb0 = (rem - mby * spr) * NMB;
            const int nmb = min(NMB, mbw - mb0);
            uint8_t *s_cur = bb + P::B_CUR;
            uint32_t *s_exp = (uint32_t *)(bb + P::B_EXP);
            uint32_t *s_costx = (uint32_t *)(bb + P::B_COSTX), *s_costy = (uint32_t *)(bb + P::B_COSTY);
            if (ptid == 0) {
                fence_proxy_async();                             // earlier generic reads of raw / cur before the async writes
                mbar_expect_tx(s_tma, S::RAW_BYTES + S::CUR_BYTES);
                tma_load_3d(s_raw, &tm_ref, B2_PAD + mb0 * 16 - R, B2_PAD + mby * 16 - R, frame, s_tma);
                tma_load_3d(s_cur, &tm_cur, B2_PAD + mb0 * 16, B2_PAD + mby * 16, frame, s_tma);
            }
            // while the TMA is in flight: predictors (one global load per macroblock), cost tables and best keys
            const size_t mb_base = ((size_t)frame * mbh + mby) * mbw + mb0;
            if (ptid < NMB) {
                int px = 0, py = 0;
                if (pmv != nullptr && ptid < nmb) { const b2_mv_t pp = pmv[mb_base + ptid]; px = pp.x; py = pp.y; }
                s_pmv[0][ptid] = px; s_pmv[1][ptid] = py;
                s_best[ptid] = 0xffffffffu;
            }
            producer_sync();
            for (int i = ptid; i < NMB * ND; i += K1P_PW * 32) {
                const int m = i / ND, d = i - m * ND;
                s_costx[i] = ((uint32_t)(lambda * b2_mvbits(4 * (d - R) - s_pmv[0][m])) << 13) + (uint32_t)d;
                s_costy[i] = ((uint32_t)(lambda * b2_mvbits(4 * (d - R) - s_pmv[1][m])) << 13) + (uint32_t)(d * ND);
            }
            mbar_wait(s_tma, (uint32_t)(k & 1));
            {   // expand: word x of a row = pixels x..x+3 (little endian), all 4 byte alignments
                constexpr int WPR = S::WIN_W / 4;
                const uint32_t *raw32 = (const uint32_t *)s_raw;
                for (int i = ptid; i < WPR * S::WIN_H; i += K1P_PW * 32) {
                    const int r = i / WPR, j = i - r * WPR;
                    const uint32_t lo = raw32[i], hi = raw32[i + 1];
                    uint4 o;
                    o.x = lo; o.y = __funnelshift_r(lo, hi, 8); o.z = __funnelshift_r(lo, hi, 16); o.w = __funnelshift_r(lo, hi, 24);
                    *(uint4 *)(s_exp + r * S::EXP_PITCH + 4 * j) = o;
                }
            }
            producer_sync();                                     // every producer thread's stores precede the release below
            if (ptid == 0) mbar_arrive(&s_full[buf]);            // release: the buffer is complete
        }
        return;
    }

    // ================= consumer warps =================
    // per-warp state is kept small (one buffer offset instead of six pointers): the sweep itself needs ~75 registers
    int cur_k = -1, nmb = 0, bo = 0;
    const int ntasks = mine * TPS;
    for (int T = warp; T < ntasks; T += NCW) {
        const int k = T / TPS, t0 = (T - k * TPS) * 32;
        if (k != cur_k) {
            cur_k = k;
            mbar_wait(&s_full[k % NBUF], (uint32_t)((k / NBUF) & 1));
            bo = P::BUF0 + (k % NBUF) * P::BUF_BYTES;
            const int sidx = (int)blockIdx.x + k * (int)gridDim.x;
            nmb = min(NMB, mbw - (sidx % spr) * NMB);
        }
        const uint8_t *s_cur = smem + bo + P::B_CUR;
        const uint32_t *s_exp = (const uint32_t *)(smem + bo + P::B_EXP);
        const uint32_t *s_costx = (const uint32_t *)(smem + bo + P::B_COSTX), *s_costy = (const uint32_t *)(smem + bo + P::B_COSTY);
        uint32_t *s_best = (uint32_t *)(smem + bo + P::B_BEST);
        const int total = nmb * NG * ND;
        if (t0 < total) {
            const int t = t0 + lane;
            const bool active = t < total;
            const int tt = active ? t : t0;               // idle lanes shadow lane 0's task
            const int m = tt / (NG * ND);
            const int rem = tt - m * (NG * ND);
            const int g = rem / ND;
            const int dxi = rem - g * ND;
            uint32_t cur[64];
            {
                const uint4 *c4 = (const uint4 *)(s_cur + m * 16);
#pragma unroll
                for (int y = 0; y < 16; y++) {
                    uint4 v = c4[y * (NMB * 16 / 16)];
                    cur[y * 4 + 0] = v.x; cur[y * 4 + 1] = v.y; cur[y * 4 + 2] = v.z; cur[y * 4 + 3] = v.w;
                }
            }
            uint32_t acc[K];
#pragma unroll
            for (int kk = 0; kk < K; kk++) acc[kk] = 0;
            const uint32_t *wp = s_exp + (g * K) * S::EXP_PITCH + m * 16 + dxi;
#pragma unroll
            for (int r = 0; r < K + 15; r++) {
                const uint32_t w0 = wp[r * S::EXP_PITCH + 0];
                const uint32_t w1 = wp[r * S::EXP_PITCH + 4];
                const uint32_t w2 = wp[r * S::EXP_PITCH + 8];
                const uint32_t w3 = wp[r * S::EXP_PITCH + 12];
#pragma unroll
                for (int kk = 0; kk < K; kk++) {
                    const int y = r - kk;
                    if (y >= 0 && y < 16) {
                        acc[kk] = vsad4_acc(w0, cur[y * 4 + 0], acc[kk]);
                        acc[kk] = vsad4_acc(w1, cur[y * 4 + 1], acc[kk]);
                        acc[kk] = vsad4_acc(w2, cur[y * 4 + 2], acc[kk]);
                        acc[kk] = vsad4_acc(w3, cur[y * 4 + 3], acc[kk]);
                    }
                }
            }
            uint32_t key = 0xffffffffu;
            if (active) {
                const uint32_t kx = s_costx[m * ND + dxi];
                const uint32_t *ky = s_costy + m * ND + g * K;
#pragma unroll
                for (int kk = 0; kk + 1 < K; kk += 2) key = __vimin3_u32(key, acc[kk] * 8192u + (kx + ky[kk]), acc[kk + 1] * 8192u + (kx + ky[kk + 1]));
                if (K & 1) key = min(key, acc[K - 1] * 8192u + (kx + ky[K - 1]));
            }
            const int m0 = __shfl_sync(0xffffffffu, m, 0);
            if (__all_sync(0xffffffffu, m == m0)) {
                const uint32_t wmin = __reduce_min_sync(0xffffffffu, key);
                if (lane == 0) atomicMin(&s_best[m0], wmin);
            } else if (active) {
                atomicMin(&s_best[m], key);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[k % NBUF]);          // release: this task's reads and atomics are done
    }
}

// ---- lossless pruning: successive elimination in front of the same sweep (engine option me_prune) ---------------------------
// |sum(cur MB) - sum(ref block)| <= SAD(cur MB, ref block) (triangle inequality), so a candidate whose
//     |C - S(dx,dy)| + mvcost(dx,dy)  >  U,     U = exact cost of any candidate already evaluated,
// cannot be the minimum and cannot tie with it either; dropping it leaves argmin AND tie-break of the exhaustive scan untouched
// (x264's own "esa" search is built on the same inequality).  Per strip:
//   1. TMA window + current tile, cost tables, expansion: as k1_me_fullpel_kernel;
//   2. the raw window is dead after the expansion, so the 16x16 block sums of the reference at the strip's (NMB-1)*16 + 2R+1
//      candidate columns x 2R+1 rows arrive by TMA into the same shared memory (K1a below computes them once per reference
//      frame: u16 plane with the geometry of the padded luma plane);
//   3. one warp per macroblock evaluates two candidates exactly -- the zero vector and the rounded predictor (the co-located
//      vector of the previous frame) -- giving U and the first best key;
//   4. every lane-task (mb, dy-group, dx) tests its K candidates against U (mvcost of the row replaced by the smallest of the
//      group: weaker, still a lower bound); the survivors are compacted (ballot -> prefix over the mask words -> list);
//   5. the unchanged register-tiled sweep runs over the survivor list only.
// What is skipped is reported, never folded into the roofline figure: `swept` counts the lane-tasks that ran (b2_engine_k1_stats).
constexpr int k1_cmax(int a, int b) { return a > b ? a : b; }
template <int R> struct K1SeaSmem {
    using S = K1Smem<R>;
    static constexpr int NMB = S::NMB, ND = S::ND, NG = K1Cfg<R>::NG;
    static constexpr int NX = (NMB - 1) * 16 + ND;               // candidate columns of a strip
    static constexpr int SUM_PITCH = (NX + 7) & ~7;              // u16 per row of the sum window (TMA rows are multiples of 16 B)
    static constexpr int SUM_BYTES = SUM_PITCH * ND * 2;
    static constexpr int NTASK = NMB * NG * ND;
    static constexpr int NWORD = (NTASK + 31) / 32;
    // region 0: raw window while it is expanded, then the block-sum window, then the survivor list (u16 lane-task ids)
    static constexpr int R0_BYTES = k1_cmax(k1_cmax(S::RAW_BYTES + 16, SUM_BYTES), NTASK * 2);
    static constexpr int OFF_R0 = 0;
    static constexpr int OFF_CUR = (R0_BYTES + 127) & ~127;
    static constexpr int OFF_EXP = OFF_CUR + S::CUR_BYTES;
    static constexpr int OFF_COSTX = OFF_EXP + S::EXP_BYTES;
    static constexpr int OFF_COSTY = OFF_COSTX + NMB * ND * 4;
    static constexpr int OFF_BEST = OFF_COSTY + NMB * ND * 4;
    static constexpr int OFF_U = OFF_BEST + NMB * 4;
    static constexpr int OFF_CSUM = OFF_U + NMB * 4;
    static constexpr int OFF_MINCY = OFF_CSUM + NMB * 4;
    static constexpr int OFF_MASK = OFF_MINCY + NMB * NG * 4;
    static constexpr int OFF_PREF = OFF_MASK + NWORD * 4;        // NWORD exclusive prefixes + the survivor count
    static constexpr int OFF_BAR = (OFF_PREF + (NWORD + 1) * 4 + 7) & ~7;
    static constexpr int TOTAL = OFF_BAR + 16 + 128;             // +128: manual alignment slack
};

// three CTAs per SM, like the exhaustive kernel: 3 x (TOTAL + 1 KB reserved per CTA) <= 228 KB
static_assert(B2_K1_MINCTAS != 3 || B2_K1_NMB != 6 || 3 * (K1SeaSmem<32>::TOTAL + 1024) <= 228 * 1024, "pruned K1 (+-32) no longer fits three CTAs per SM");
static_assert(K1SeaSmem<32>::NTASK < 65536 && K1SeaSmem<16>::NTASK < 65536, "survivor list entries are u16");

template <int R, int NTHREADS>
__global__ void __launch_bounds__(NTHREADS, B2_K1_MINCTAS)
k1_me_fullpel_sea_kernel(const __grid_constant__ CUtensorMap tm_cur, const __grid_constant__ CUtensorMap tm_ref,
                         const __grid_constant__ CUtensorMap tm_sum, int mbw, int mbh, const b2_mv_t *__restrict__ pmv, int lambda,
                         b2_mv_t *__restrict__ mv_out, uint32_t *__restrict__ cost_out, unsigned long long *__restrict__ swept)
{
    using S = K1Smem<R>;
    using Q = K1SeaSmem<R>;
    constexpr int NWARPS = NTHREADS / 32;
    constexpr int K = K1Cfg<R>::K, NG = K1Cfg<R>::NG, ND = S::ND, NMB = K1Cfg<R>::NMB;

    extern __shared__ uint8_t smem_raw_[];
    uint8_t *smem = smem_raw_ + ((128u - (smem_u32(smem_raw_) & 127u)) & 127u);
    uint8_t *s_raw = smem + Q::OFF_R0;
    const uint16_t *s_sum = (const uint16_t *)(smem + Q::OFF_R0);
    uint16_t *s_list = (uint16_t *)(smem + Q::OFF_R0);
    uint8_t *s_cur = smem + Q::OFF_CUR;
    uint32_t *s_exp = (uint32_t *)(smem + Q::OFF_EXP);
    uint32_t *s_costx = (uint32_t *)(smem + Q::OFF_COSTX);
    uint32_t *s_costy = (uint32_t *)(smem + Q::OFF_COSTY);
    uint32_t *s_best = (uint32_t *)(smem + Q::OFF_BEST);
    int *s_u = (int *)(smem + Q::OFF_U);
    int *s_csum = (int *)(smem + Q::OFF_CSUM);
    int *s_mincy = (int *)(smem + Q::OFF_MINCY);
    uint32_t *s_mask = (uint32_t *)(smem + Q::OFF_MASK);
    uint32_t *s_pref = (uint32_t *)(smem + Q::OFF_PREF);
    uint64_t *s_bar = (uint64_t *)(smem + Q::OFF_BAR);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mb0 = blockIdx.x * NMB, mby = blockIdx.y, frame = blockIdx.z;
    const int nmb = min(NMB, mbw - mb0);

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&s_bar[0], S::RAW_BYTES + S::CUR_BYTES);
        tma_load_3d(s_raw, &tm_ref, B2_PAD + mb0 * 16 - R, B2_PAD + mby * 16 - R, frame, &s_bar[0]);
        tma_load_3d(s_cur, &tm_cur, B2_PAD + mb0 * 16, B2_PAD + mby * 16, frame, &s_bar[0]);
    }

    const size_t mb_base = ((size_t)frame * mbh + mby) * mbw + mb0;
    for (int i = tid; i < NMB * ND; i += NTHREADS) {
        int m = i / ND, d = i - m * ND;
        int px = 0, py = 0;
        if (pmv != nullptr && m < nmb) { b2_mv_t p = pmv[mb_base + m]; px = p.x; py = p.y; }
        s_costx[i] = ((uint32_t)(lambda * b2_mvbits(4 * (d - R) - px)) << 13) + (uint32_t)d;
        s_costy[i] = ((uint32_t)(lambda * b2_mvbits(4 * (d - R) - py)) << 13) + (uint32_t)(d * ND);
    }

    mbar_wait(&s_bar[0], 0);
    {
        constexpr int WPR = S::WIN_W / 4;
        const uint32_t *raw32 = (const uint32_t *)s_raw;
        for (int i = tid; i < WPR * S::WIN_H; i += NTHREADS) {
            int r = i / WPR, j = i - r * WPR;
            uint32_t lo = raw32[i], hi = raw32[i + 1];
            uint4 o;
            o.x = lo;
            o.y = __funnelshift_r(lo, hi, 8);
            o.z = __funnelshift_r(lo, hi, 16);
            o.w = __funnelshift_r(lo, hi, 24);
            *(uint4 *)(s_exp + r * S::EXP_PITCH + 4 * j) = o;
        }
    }
    fence_proxy_async();                  // this thread's generic reads of the raw window precede the TMA write into the same bytes
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&s_bar[1], Q::SUM_BYTES);
        tma_load_3d(s_raw, &tm_sum, B2_PAD + mb0 * 16 - R, B2_PAD + mby * 16 - R, frame, &s_bar[1]);
    }

    // thresholds while the sums are in flight: exact keys of the zero vector and of the rounded predictor
    for (int m = warp; m < nmb; m += NWARPS) {
        int px = 0, py = 0;
        if (pmv != nullptr) { b2_mv_t p = pmv[mb_base + m]; px = p.x; py = p.y; }
        const int cdx = min(max((px + 2) >> 2, -R), R) + R, cdy = min(max((py + 2) >> 2, -R), R) + R;
        uint32_t sa = 0, sb = 0, cs = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int w = lane + 32 * h, y = w >> 2, c = w & 3;
            const uint32_t cw = ((const uint32_t *)s_cur)[y * (NMB * 4) + m * 4 + c];
            sa = vsad4_acc(s_exp[(R + y) * S::EXP_PITCH + m * 16 + R + 4 * c], cw, sa);
            sb = vsad4_acc(s_exp[(cdy + y) * S::EXP_PITCH + m * 16 + cdx + 4 * c], cw, sb);
            cs = __dp4a(cw, 0x01010101u, cs);
        }
        sa = __reduce_add_sync(0xffffffffu, sa);
        sb = __reduce_add_sync(0xffffffffu, sb);
        cs = __reduce_add_sync(0xffffffffu, cs);
        if (lane == 0) {
            const uint32_t ka = sa * 8192u + s_costx[m * ND + R] + s_costy[m * ND + R];
            const uint32_t kb = sb * 8192u + s_costx[m * ND + cdx] + s_costy[m * ND + cdy];
            const uint32_t kk = min(ka, kb);
            s_best[m] = kk;
            s_u[m] = (int)(kk >> 13);
            s_csum[m] = (int)cs;
        }
    }
    for (int i = tid; i < NMB * NG; i += NTHREADS) {
        const int m = i / NG, g = i - m * NG;
        uint32_t mn = 0xffffffffu;
        for (int k = 0; k < K; k++) mn = min(mn, s_costy[m * ND + g * K + k] >> 13);
        s_mincy[i] = (int)mn;
    }
    __syncthreads();
    mbar_wait(&s_bar[1], 0);

    // bound test: lane-task id = (round * NWARPS + warp) * 32 + lane, so one warp round fills exactly one mask word
    const int total = nmb * NG * ND;
    for (int id0 = warp * 32; id0 < total; id0 += NWARPS * 32) {
        const int id = id0 + lane;
        bool alive = false;
        if (id < total) {
            const int m = id / (NG * ND);
            const int rem = id - m * (NG * ND);
            const int g = rem / ND;
            const int dxi = rem - g * ND;
            const int t = s_u[m] - (int)(s_costx[m * ND + dxi] >> 13) - s_mincy[m * NG + g];
            if (t >= 0) {
                const uint16_t *sp = s_sum + (g * K) * Q::SUM_PITCH + m * 16 + dxi;
                const int c = s_csum[m];
#pragma unroll
                for (int k = 0; k < K; k++) alive |= abs((int)sp[k * Q::SUM_PITCH] - c) <= t;
            }
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) s_mask[id0 >> 5] = mask;
    }
    __syncthreads();                      // all reads of the sum window are done: region 0 becomes the survivor list
    const int nword = (total + 31) >> 5;
    if (warp == 0) {
        uint32_t run = 0;
        for (int base = 0; base < nword; base += 32) {
            const int w = base + lane;
            const uint32_t c = w < nword ? (uint32_t)__popc(s_mask[w]) : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (w < nword) s_pref[w] = run + inc - c;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_pref[Q::NWORD] = run;
    }
    __syncthreads();
    const int count = (int)s_pref[Q::NWORD];
    for (int id0 = warp * 32; id0 < total; id0 += NWARPS * 32) {
        const uint32_t mask = s_mask[id0 >> 5];
        if ((mask >> lane) & 1u) s_list[s_pref[id0 >> 5] + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)(id0 + lane);
    }
    __syncthreads();

    for (int t0 = warp * 32; t0 < count; t0 += NWARPS * 32) {
        const bool active = t0 + lane < count;
        const int tt = s_list[active ? t0 + lane : t0];       // idle lanes shadow lane 0's task
        const int m = tt / (NG * ND);
        const int rem = tt - m * (NG * ND);
        const int g = rem / ND;
        const int dxi = rem - g * ND;

        uint32_t cur[64];
        {
            const uint4 *c4 = (const uint4 *)(s_cur + m * 16);
#pragma unroll
            for (int y = 0; y < 16; y++) {
                uint4 v = c4[y * (NMB * 16 / 16)];
                cur[y * 4 + 0] = v.x; cur[y * 4 + 1] = v.y; cur[y * 4 + 2] = v.z; cur[y * 4 + 3] = v.w;
            }
        }
        uint32_t acc[K];
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = 0;
        const uint32_t *wp = s_exp + (g * K) * S::EXP_PITCH + m * 16 + dxi;
#pragma unroll
        for (int r = 0; r < K + 15; r++) {
            const uint32_t w0 = wp[r * S::EXP_PITCH + 0];
            const uint32_t w1 = wp[r * S::EXP_PITCH + 4];
            const uint32_t w2 = wp[r * S::EXP_PITCH + 8];
            const uint32_t w3 = wp[r * S::EXP_PITCH + 12];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int y = r - k;
                if (y >= 0 && y < 16) {
                    acc[k] = vsad4_acc(w0, cur[y * 4 + 0], acc[k]);
                    acc[k] = vsad4_acc(w1, cur[y * 4 + 1], acc[k]);
                    acc[k] = vsad4_acc(w2, cur[y * 4 + 2], acc[k]);
                    acc[k] = vsad4_acc(w3, cur[y * 4 + 3], acc[k]);
                }
            }
        }
        uint32_t key = 0xffffffffu;
        if (active) {
            const uint32_t kx = s_costx[m * ND + dxi];
            const uint32_t *ky = s_costy + m * ND + g * K;
#pragma unroll
            for (int k = 0; k + 1 < K; k += 2) key = __vimin3_u32(key, acc[k] * 8192u + (kx + ky[k]), acc[k + 1] * 8192u + (kx + ky[k + 1]));
            if (K & 1) key = min(key, acc[K - 1] * 8192u + (kx + ky[K - 1]));
        }
        const int m0 = __shfl_sync(0xffffffffu, m, 0);
        if (__all_sync(0xffffffffu, m == m0)) {
            const uint32_t wmin = __reduce_min_sync(0xffffffffu, key);
            if (lane == 0) atomicMin(&s_best[m0], wmin);
        } else if (active) {
            atomicMin(&s_best[m], key);
        }
    }
    __syncthreads();

    if (tid < nmb) {
        const uint32_t key = s_best[tid];
        const int idx = (int)(key & 8191u);
        const int dyi = idx / ND, dxi = idx - dyi * ND;
        b2_mv_t mv;
        mv.x = (int16_t)(dxi - R);
        mv.y = (int16_t)(dyi - R);
        mv_out[mb_base + tid] = mv;
        cost_out[mb_base + tid] = key >> 13;
    }
    if (tid == 0 && swept != nullptr) atomicAdd(swept, (unsigned long long)count);
}

// K1a: 16x16 block sums of a padded luma plane, sum[Y][X] = sum of plane[Y..Y+15][X..X+15] (u16: at most 65,280), for every
// position whose block lies inside the allocation; same [n][rows][pitch] geometry as the plane.  HBM-bound and tiny next to the
// search (1 B read + 2 B written per pixel): one CTA per 128 x 32 tile, vertical 16-sums by a sliding column walk, then the
// horizontal 16-sum as two levels of 4.
constexpr int K1A_TW = 128, K1A_TH = 32;
__global__ void __launch_bounds__(256) k1a_block_sums_kernel(const uint8_t *__restrict__ planes, int pitch, int rows, size_t plane_stride,
                                                             uint16_t *__restrict__ sums)
{
    __shared__ uint32_t s_in[K1A_TH + 15][(K1A_TW + 16) / 4];
    __shared__ uint16_t s_v[K1A_TH][K1A_TW + 16];
    __shared__ uint16_t s_q[K1A_TH][K1A_TW + 16];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * K1A_TW, y0 = blockIdx.y * K1A_TH;
    const uint8_t *plane = planes + (size_t)blockIdx.z * plane_stride;
    uint16_t *out = sums + (size_t)blockIdx.z * plane_stride;           // same element stride: the sum plane mirrors the luma plane
    constexpr int WPR = (K1A_TW + 16) / 4;
    for (int i = tid; i < (K1A_TH + 15) * WPR; i += 256) {
        const int r = i / WPR, w = i - r * WPR;
        const int yy = min(y0 + r, rows - 1), xx = x0 + 4 * w;
        s_in[r][w] = xx + 3 < pitch ? *(const uint32_t *)(plane + (size_t)yy * pitch + xx) : 0u;
    }
    __syncthreads();
    if (tid < K1A_TW + 16) {
        const uint8_t *col = (const uint8_t *)&s_in[0][0] + tid;
        constexpr int P = WPR * 4;
        int acc = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) acc += col[j * P];
        s_v[0][tid] = (uint16_t)acc;
        for (int y = 1; y < K1A_TH; y++) {
            acc += col[(y + 15) * P] - col[(y - 1) * P];
            s_v[y][tid] = (uint16_t)acc;
        }
    }
    __syncthreads();
    for (int i = tid; i < K1A_TH * (K1A_TW + 12); i += 256) {
        const int y = i / (K1A_TW + 12), x = i - y * (K1A_TW + 12);
        s_q[y][x] = (uint16_t)(s_v[y][x] + s_v[y][x + 1] + s_v[y][x + 2] + s_v[y][x + 3]);
    }
    __syncthreads();
    for (int i = tid; i < K1A_TH * K1A_TW; i += 256) {
        const int y = i / K1A_TW, x = i - y * K1A_TW;
        const int Y = y0 + y, X = x0 + x;
        if (Y < rows && X < pitch) {
            const bool valid = Y + 16 <= rows && X + 16 <= pitch;
            out[(size_t)Y * pitch + X] = valid ? (uint16_t)(s_q[y][x] + s_q[y][x + 4] + s_q[y][x + 8] + s_q[y][x + 12]) : (uint16_t)0;
        }
    }
}

#ifndef B2_K1P_WARPS
#define B2_K1P_WARPS 16      // 16 consumer warps (92 registers, no spills) + 4 producer warps; 20 consumers fit only at 80 registers
#endif
template <int R>
int launch_k1_persistent(const CUtensorMap &tm_cur, const CUtensorMap &tm_ref, int mbw, int mbh, int nframes, const b2_mv_t *pmv,
                         int lambda, b2_mv_t *mv_out, uint32_t *cost_out, cudaStream_t st)
{
    static std::atomic<int> nsm[64];                          // SM count per device, 0 = not looked up yet
    int dev = 0;
    B2_CUDA_OK(cudaGetDevice(&dev));
    int sms = (dev >= 0 && dev < 64) ? nsm[dev].load(std::memory_order_acquire) : 0;
    if (!sms) {
        B2_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        B2_CUDA_OK(cudaFuncSetAttribute(k1_me_fullpel_persistent_kernel<R, B2_K1P_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        K1PSmem<R>::TOTAL));
        if (dev >= 0 && dev < 64) nsm[dev].store(sms, std::memory_order_release);
    }
    constexpr int NMB = K1Cfg<R>::NMB;
    const int nstrips = ((mbw + NMB - 1) / NMB) * mbh * nframes;
    const int grid = nstrips < sms ? nstrips : sms;
    k1_me_fullpel_persistent_kernel<R, B2_K1P_WARPS><<<grid, (B2_K1P_WARPS + K1P_PW) * 32, K1PSmem<R>::TOTAL, st>>>(
        tm_cur, tm_ref, mbw, mbh, nframes, pmv, lambda, mv_out, cost_out);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

// B2_K1_PERSISTENT=1 selects the persistent, pipelined kernel for the single-vector search.  Measured on the B200 (1080p, +-32,
// 64 frames per launch): 0.889 of the VABSDIFF4 peak with 16 consumer + 4 producer warps and three buffers (0.866 with two
// buffers and two producer warps, 0.861 with 20 consumer warps at 80 registers) against 0.888 for one strip per CTA x 3 resident
// CTAs, and 5,333 vs 5,361 frames/s for the whole step: both forms sit at the same ceiling, so the simpler one stays the default.
int k1_persistent()
{
    static const int v = [] { const char *e = getenv("B2_K1_PERSISTENT"); return e ? atoi(e) : 0; }();
    return v;
}

template <int R, int NT, bool PART>
int launch_k1p(const CUtensorMap &tm_cur, const CUtensorMap &tm_ref, int mbw, int mbh, int nframes, const b2_mv_t *pmv, int lambda,
               b2_mv_t *mv_out, uint32_t *cost_out, b2_mv_t *mv9_out, uint32_t *cost9_out, cudaStream_t st)
{
    // function attributes are per device: one process may drive several GPUs, each from its own host thread
    static std::atomic<bool> attr_set[64];
    int dev = 0;
    B2_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
        B2_CUDA_OK(cudaFuncSetAttribute(k1_me_fullpel_kernel<R, NT, PART>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        K1Smem<R>::TOTAL + B2_K1_SMEM_PAD));
        if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
    }
    constexpr int NMB = K1Cfg<R>::NMB;
    dim3 grid((mbw + NMB - 1) / NMB, mbh, nframes);
    k1_me_fullpel_kernel<R, NT, PART><<<grid, NT, K1Smem<R>::TOTAL + B2_K1_SMEM_PAD, st>>>(tm_cur, tm_ref, mbw, mbh, pmv, lambda, mv_out, cost_out,
                                                                            mv9_out, cost9_out);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}
template <int R, int NT>
int launch_k1(const CUtensorMap &tm_cur, const CUtensorMap &tm_ref, int mbw, int mbh, int nframes, const b2_mv_t *pmv, int lambda,
              b2_mv_t *mv_out, uint32_t *cost_out, b2_mv_t *mv9_out, uint32_t *cost9_out, cudaStream_t st)
{
    if (mv9_out) return launch_k1p<R, NT, true>(tm_cur, tm_ref, mbw, mbh, nframes, pmv, lambda, mv_out, cost_out, mv9_out, cost9_out, st);
    return launch_k1p<R, NT, false>(tm_cur, tm_ref, mbw, mbh, nframes, pmv, lambda, mv_out, cost_out, nullptr, nullptr, st);
}

template <int R, int NT>
int launch_k1_sea(const CUtensorMap &tm_cur, const CUtensorMap &tm_ref, const CUtensorMap &tm_sum, int mbw, int mbh, int nframes,
                  const b2_mv_t *pmv, int lambda, b2_mv_t *mv_out, uint32_t *cost_out, unsigned long long *swept, cudaStream_t st)
{
    static std::atomic<bool> attr_set[64];
    int dev = 0;
    B2_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
        B2_CUDA_OK(cudaFuncSetAttribute(k1_me_fullpel_sea_kernel<R, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1SeaSmem<R>::TOTAL));
        if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
    }
    constexpr int NMB = K1Cfg<R>::NMB;
    dim3 grid((mbw + NMB - 1) / NMB, mbh, nframes);
    k1_me_fullpel_sea_kernel<R, NT><<<grid, NT, K1SeaSmem<R>::TOTAL, st>>>(tm_cur, tm_ref, tm_sum, mbw, mbh, pmv, lambda, mv_out, cost_out, swept);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

// threads per CTA: 82 (R=32) / 41 (R=16) warp-tasks per full strip should divide evenly over the warps
int k1_threads()
{
    static const int nt = [] {                          // thread-safe one-time initialisation
        const char *e = getenv("B2_K1_THREADS");
        const int v = e ? atoi(e) : 256;
        return (v == 192 || v == 256 || v == 320 || v == 384) ? v : 256;
    }();
    return nt;
}

}  // namespace

extern "C" int b2_k1_strip_mbs(int R) { return R == 16 ? K1Cfg<16>::NMB : K1Cfg<32>::NMB; }

// window/current-tile box sizes needed to build the tensor maps
extern "C" int b2_k1_window_box(int R, int *bw, int *bh)
{
    if (R != 16 && R != 32) return -1;
    *bw = b2_k1_strip_mbs(R) * 16 + 2 * R;
    *bh = 16 + 2 * R;
    return 0;
}

// d_* are device pointers; the tensor maps describe [nframes][rows][pitch] padded luma planes
// with boxes {128+2R, 16+2R, 1} (ref) and {128, 16, 1} (cur).
int b2_launch_me_fullpel(int R, const CUtensorMap *tm_cur, const CUtensorMap *tm_ref, int mbw, int mbh,
                         int nframes, const b2_mv_t *d_pmv, int lambda, b2_mv_t *d_mv, uint32_t *d_cost,
                         b2_mv_t *d_mv9, uint32_t *d_cost9, cudaStream_t st)
{
#define K1_DISPATCH(RR)                                                                                              \
    if (!d_mv9 && k1_persistent()) return launch_k1_persistent<RR>(*tm_cur, *tm_ref, mbw, mbh, nframes, d_pmv, lambda, d_mv, d_cost, st); \
    switch (k1_threads()) {                                                                                          \
    case 192: return launch_k1<RR, 192>(*tm_cur, *tm_ref, mbw, mbh, nframes, d_pmv, lambda, d_mv, d_cost, d_mv9, d_cost9, st);       \
    case 320: return launch_k1<RR, 320>(*tm_cur, *tm_ref, mbw, mbh, nframes, d_pmv, lambda, d_mv, d_cost, d_mv9, d_cost9, st);       \
    case 384: return launch_k1<RR, 384>(*tm_cur, *tm_ref, mbw, mbh, nframes, d_pmv, lambda, d_mv, d_cost, d_mv9, d_cost9, st);       \
    default: return launch_k1<RR, 256>(*tm_cur, *tm_ref, mbw, mbh, nframes, d_pmv, lambda, d_mv, d_cost, d_mv9, d_cost9, st);        \
    }
    switch (R) {
    case 32: K1_DISPATCH(32)
    case 16: K1_DISPATCH(16)
    default:
        fprintf(stderr, "b2enc: merange %d not supported (16 or 32)\n", R);
        return -1;
    }
}

// ---- pruned search (me_prune): block sums of the reference (K1a), then K1 with successive elimination ------------------------
// box of the block-sum window a strip fetches: {((NMB-1)*16 + 2R+1 rounded up to 8) u16, 2R+1 rows, 1}
extern "C" int b2_k1_sum_box(int R, int *bw, int *bh)
{
    if (R == 32) { *bw = K1SeaSmem<32>::SUM_PITCH; *bh = K1SeaSmem<32>::ND; return 0; }
    if (R == 16) { *bw = K1SeaSmem<16>::SUM_PITCH; *bh = K1SeaSmem<16>::ND; return 0; }
    return -1;
}

int b2_launch_block_sums(const uint8_t *d_planes, int pitch, int rows, int nplanes, uint16_t *d_sums, cudaStream_t st)
{
    dim3 grid((pitch + K1A_TW - 1) / K1A_TW, (rows + K1A_TH - 1) / K1A_TH, nplanes);
    k1a_block_sums_kernel<<<grid, 256, 0, st>>>(d_planes, pitch, rows, (size_t)pitch * rows, d_sums);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

// lane-tasks (mb, dy-group, dx) of an exhaustive launch: what `swept` is compared with
extern "C" long long b2_k1_lane_tasks(int R, int mbw, int mbh, int nframes)
{
    const long long per_mb = R == 16 ? (long long)K1Cfg<16>::NG * (2 * 16 + 1) : (long long)K1Cfg<32>::NG * (2 * 32 + 1);
    return per_mb * mbw * mbh * nframes;
}

int b2_launch_me_fullpel_pruned(int R, const CUtensorMap *tm_cur, const CUtensorMap *tm_ref, const CUtensorMap *tm_sum, int mbw, int mbh,
                                int nframes, const b2_mv_t *d_pmv, int lambda, b2_mv_t *d_mv, uint32_t *d_cost,
                                unsigned long long *d_swept, cudaStream_t st)
{
#define K1S_DISPATCH(RR)                                                                                                          \
    switch (k1_threads()) {                                                                                                       \
    case 192: return launch_k1_sea<RR, 192>(*tm_cur, *tm_ref, *tm_sum, mbw, mbh, nframes, d_pmv, lambda, d_mv, d_cost, d_swept, st); \
    case 320: return launch_k1_sea<RR, 320>(*tm_cur, *tm_ref, *tm_sum, mbw, mbh, nframes, d_pmv, lambda, d_mv, d_cost, d_swept, st); \
    case 384: return launch_k1_sea<RR, 384>(*tm_cur, *tm_ref, *tm_sum, mbw, mbh, nframes, d_pmv, lambda, d_mv, d_cost, d_swept, st); \
    default: return launch_k1_sea<RR, 256>(*tm_cur, *tm_ref, *tm_sum, mbw, mbh, nframes, d_pmv, lambda, d_mv, d_cost, d_swept, st);  \
    }
    switch (R) {
    case 32: K1S_DISPATCH(32)
    case 16: K1S_DISPATCH(16)
    default:
        fprintf(stderr, "b2enc: merange %d not supported (16 or 32)\n", R);
        return -1;
    }
}
