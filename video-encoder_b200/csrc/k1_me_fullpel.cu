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
//  * Lanes of a warp are 32 consecutive dx -> consecutive words -> no bank conflicts.
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
#define B2_K1_NMB16 12      // +-16: the window is small (33 x 3 dy groups = 99 lane-tasks per MB), a wider strip fills the 8 warps' rounds
#endif

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
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const uint32_t tl = top[k] & 0xffffu, tr = top[k] >> 16, bl = accl[k], br = accr[k];
                        const uint32_t kc = kx + ky[k];
                        const uint32_t st = tl + tr, sb = bl + br;
                        best[0] = min(best[0], (st + sb) * 8192u + kc);
                        best[1] = min(best[1], st * 8192u + kc);
                        best[2] = min(best[2], sb * 8192u + kc);
                        best[3] = min(best[3], (tl + bl) * 8192u + kc);
                        best[4] = min(best[4], (tr + br) * 8192u + kc);
                        best[5] = min(best[5], tl * 8192u + kc);
                        best[6] = min(best[6], tr * 8192u + kc);
                        best[7] = min(best[7], bl * 8192u + kc);
                        best[8] = min(best[8], br * 8192u + kc);
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
#pragma unroll
            for (int k = 0; k < K; k++) key = min(key, acc[k] * 8192u + (kx + ky[k]));
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
