// k5_recon.cu -- K4/K5/K7: mode decision, residual -> 4x4 DCT -> quant -> dequant -> IDCT ->
// reconstruction, for inter macroblocks (K5, fully parallel) and intra macroblocks (K7, wavefront).
//
// Replaces x264's macroblock encode + reconstruct (behind x264_encoder_encode, av_encode.c:970);
// bit-exact against oracle/b2o_encode.c (b2o_recon_inter_mb / b2o_recon_intra_mb) and therefore,
// through the decoder drift test, identical to what libavcodec's H.264 decoder reconstructs.
//
// Work split inside one warp per macroblock: lanes 0-15 own the sixteen luma 4x4 blocks (z order),
// lanes 16-23 the eight chroma 4x4 blocks (U0-3, V0-3); DC Hadamards run across lanes by shuffle.
// K7 keeps the serial dependency of intra prediction on reconstructed neighbours: one CTA per
// frame walks the anti-diagonals d = mbx + 2*mby (left, top, top-left and top-right MBs are all
// on earlier diagonals); inside an I4x4 macroblock the sixteen blocks run as a 10-step wavefront.
// Bound: HBM/latency (read cur + pred 3*W*H, write recon 1.5*W*H + 832 B/MB of levels).
#include "b2_mbcode.cuh"

namespace {

using namespace b2;

// ---- K5: decision + inter reconstruction ---------------------------------------------------------
constexpr int K5_WARPS = 4;

__global__ void __launch_bounds__(K5_WARPS * 32)
k5_decide_inter_kernel(FramePlanes fp, int mbw, int mbh, int nmb_total, int is_p, int do_intra, int qp,
                       const b2_mv_t *__restrict__ mvq, const uint32_t *__restrict__ cost_inter,
                       const uint32_t *__restrict__ c16, const uint32_t *__restrict__ c4,
                       b2_mbinfo_t *__restrict__ info, b2_mbcoef_t *__restrict__ coef, b2_mv_t *__restrict__ prev_mv_out,
                       const uint8_t *__restrict__ pred_y)
{
    const int lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * K5_WARPS + (threadIdx.x >> 5);
    if (mbi >= nmb_total) return;
    const int per_frame = mbw * mbh;
    const int frame = mbi / per_frame, r = mbi - frame * per_frame;
    const int mby = r / mbw, mbx = r - mby * mbw;

    // K4: decision (identical on all lanes)
    uint32_t ci = 0xffffffffu; int it = B2_MB_I16x16;
    if (do_intra) { ci = c16[mbi]; const uint32_t c = c4[mbi]; if (c < ci) { ci = c; it = B2_MB_I4x4; } }
    const uint32_t cinter = is_p ? cost_inter[mbi] : 0u;
    const bool inter = is_p && !(do_intra && ci < cinter);
    b2_mv_t mv = {0, 0};
    if (inter) mv = mvq[mbi];
    if (lane == 0) {
        info[mbi].mvx = mv.x; info[mbi].mvy = mv.y;
        info[mbi].mb_type = (uint8_t)(inter ? B2_MB_P16x16 : it);
        info[mbi].cost = inter ? cinter : ci;
        prev_mv_out[mbi] = mv;
    }
    if (!inter) return;                                   // intra MBs are reconstructed by K7

    b2_mbcoef_t *cf = &coef[mbi];
    int flags = 0;
    if (lane < 16) {
        const QParams q = make_qparams(qp, false);
        const int bx = blk_x(lane) * 4, by = blk_y(lane) * 4;
        const size_t off = (size_t)(B2_PAD + mby * 16 + by) * fp.pitch + B2_PAD + mbx * 16 + bx;
        int src[16], pred[16];
        load_src4x4(fp.cur[0] + frame * fp.stride_y + off, fp.pitch, src);
        // motion-compensated prediction of the chosen MV, produced by K2 from its half-sample planes
        load_src4x4(pred_y + (size_t)mbi * 256 + by * 16 + bx, 16, pred);
        flags = code_luma4x4(src, pred, q, cf->blk[lane], fp.rec[0] + frame * fp.stride_y + off, fp.pitch) ? 1 : 0;
    }
    {
        const bool act = lane >= 16 && lane < 24;
        const int pl = (lane >> 2) & 1, k = lane & 3;
        const int qpc = chroma_qp(qp);
        const QParams q = make_qparams(qpc, false);
        const int cbx = (k & 1) * 4, cby = (k >> 1) * 4;
        const size_t off = (size_t)(B2_PADC + mby * 8 + cby) * fp.pitchc + B2_PADC + mbx * 8 + cbx;
        int src[16], pred[16];
        if (act) {
            load_src4x4(fp.cur[1 + pl] + frame * fp.stride_c + off, fp.pitchc, src);
            const uint8_t *rp = fp.ref[1 + pl] + frame * fp.stride_c + off + (ptrdiff_t)(mv.y >> 3) * fp.pitchc + (mv.x >> 3);
            const int fx = mv.x & 7, fy = mv.y & 7;
#pragma unroll
            for (int y = 0; y < 4; y++)
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    const uint8_t *qq = rp + (size_t)y * fp.pitchc + x;
                    pred[y * 4 + x] = ((8 - fx) * (8 - fy) * qq[0] + fx * (8 - fy) * qq[1] + (8 - fx) * fy * qq[fp.pitchc] +
                                       fx * fy * qq[fp.pitchc + 1] + 32) >> 6;
                }
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) src[i] = pred[i] = 0;
        }
        const int cfl = code_chroma4x4(lane, act, src, pred, q, qpc, cf, fp.rec[1 + pl] + frame * fp.stride_c + off, fp.pitchc);
        if (act) flags = cfl;
    }
    if (lane == 24) {                                     // unused parts of the level record
        uint4 z4 = make_uint4(0, 0, 0, 0);
        ((uint4 *)cf->blk[24])[0] = z4; ((uint4 *)cf->blk[24])[1] = z4;
        ((uint4 *)cf->blk[25])[1] = z4;
    }
    const uint32_t mask = gather_mask(lane, flags, false);
    if (lane == 0) {
        info[mbi].cbp = (uint8_t)cbp_from_mask(B2_MB_P16x16, mask);
        info[mbi].nnz_mask = mask;
    }
}

}  // namespace

int b2_launch_decide_inter(const uint8_t *const cur[3], const uint8_t *const ref[3], uint8_t *const rec[3], int pitch,
                           int pitchc, size_t stride_y, size_t stride_c, int mbw, int mbh, int nframes, int is_p,
                           int do_intra, int qp, const b2_mv_t *d_mvq, const uint32_t *d_cost_inter, const uint32_t *d_c16,
                           const uint32_t *d_c4, b2_mbinfo_t *d_info, b2_mbcoef_t *d_coef, b2_mv_t *d_prev_mv,
                           const uint8_t *d_pred_y, cudaStream_t st)
{
    FramePlanes fp;
    for (int i = 0; i < 3; i++) { fp.cur[i] = cur[i]; fp.ref[i] = ref ? ref[i] : nullptr; fp.rec[i] = rec[i]; }
    fp.pitch = pitch; fp.pitchc = pitchc; fp.stride_y = stride_y; fp.stride_c = stride_c;
    const int nmb = mbw * mbh * nframes;
    k5_decide_inter_kernel<<<(nmb + K5_WARPS - 1) / K5_WARPS, K5_WARPS * 32, 0, st>>>(
        fp, mbw, mbh, nmb, is_p, do_intra, qp, d_mvq, d_cost_inter, d_c16, d_c4, d_info, d_coef, d_prev_mv, d_pred_y);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

