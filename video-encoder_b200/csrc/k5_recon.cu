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

// 8x8 transform of one inter macroblock (row N1): lanes 4q..4q+3 cooperate on 8x8 block q through a per-warp
// 16x16 int tile in shared memory (column pass, row pass + quantiser, zig-zag gather, inverse rows, inverse columns).
// ALL 32 lanes must call; lanes >= 16 only take part in the warp barriers and shuffles.  Returns this lane's nnz flag
// (bit of nnz_mask: the k-th interleaved quarter of the 8x8 block, k = lane & 3).
__device__ __forceinline__ int code_luma8x8_quad(int lane, int *sm, const int src[16], const int pred[16], int qp, int16_t *lev,
                                                 uint8_t *rec, int rpitch)
{
    const bool act = lane < 16;
    const int q = (lane >> 2) & 3, r = lane & 3, qx = (q & 1) * 8, qy = (q >> 1) * 8;
    const int bx = blk_x(lane & 15) * 4, by = blk_y(lane & 15) * 4;
    if (act) {
#pragma unroll
        for (int i = 0; i < 16; i++) sm[(by + (i >> 2)) * 16 + bx + (i & 3)] = src[i] - pred[i];
    }
    __syncwarp();
    if (act) {                                            // forward, columns first (x264's order)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            int a[8], o[8];
            const int x = qx + 2 * r + j;
#pragma unroll
            for (int y = 0; y < 8; y++) a[y] = sm[(qy + y) * 16 + x];
            fdct8_1d(a, o);
#pragma unroll
            for (int y = 0; y < 8; y++) sm[(qy + y) * 16 + x] = o[y];
        }
    }
    __syncwarp();
    int z[2][8];
    const int qbits = 16 + qp / 6, f = ((1 << qbits) * 11) >> 6, sh = qp / 6, rem = qp % 6;
    if (act) {                                            // rows, then the dead-zone quantiser on the two rows this lane owns
#pragma unroll
        for (int j = 0; j < 2; j++) {
            int a[8], o[8];
            const int y = 2 * r + j;
#pragma unroll
            for (int x = 0; x < 8; x++) a[x] = sm[(qy + y) * 16 + qx + x];
            fdct8_1d(a, o);
#pragma unroll
            for (int x = 0; x < 8; x++) {
                const uint32_t mf = c_quant8_mf[rem][c_cls8[(y & 3) * 4 + (x & 3)]];
                const int v = (int)(((uint32_t)abs(o[x]) * mf + (uint32_t)f) >> qbits);
                z[j][x] = o[x] < 0 ? -v : v;
            }
        }
    }
    __syncwarp();
    if (act) {
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int x = 0; x < 8; x++) sm[(qy + 2 * r + j) * 16 + qx + x] = z[j][x];
    }
    __syncwarp();
    int mask4 = 0;
    if (act) {                                            // levels 16r..16r+15 of the 8x8 zig-zag scan -> blk[4q + r]
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int idx = c_zigzag8[16 * r + i];
            const int v = sm[(qy + (idx >> 3)) * 16 + qx + (idx & 7)];
            if (v) mask4 |= 1 << (i & 3);
            if (i & 1) pk[i >> 1] |= (uint32_t)(uint16_t)(int16_t)v << 16;
            else pk[i >> 1] = (uint32_t)(uint16_t)(int16_t)v;
        }
        uint4 *d4 = (uint4 *)lev;
        d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    mask4 |= __shfl_xor_sync(0xffffffffu, mask4, 1);
    mask4 |= __shfl_xor_sync(0xffffffffu, mask4, 2);
    __syncwarp();
    if (act && mask4) {                                   // normative scaling + inverse rows (8.5.13)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            int a[8], o[8];
            const int y = 2 * r + j;
#pragma unroll
            for (int x = 0; x < 8; x++) {
                const int ls = 16 * c_dequant8_v[rem][c_cls8[(y & 3) * 4 + (x & 3)]];
                a[x] = sh >= 6 ? (z[j][x] * ls) * (1 << (sh - 6)) : (z[j][x] * ls + (1 << (5 - sh))) >> (6 - sh);
            }
            idct8_1d(a, o);
#pragma unroll
            for (int x = 0; x < 8; x++) sm[(qy + y) * 16 + qx + x] = o[x];
        }
    }
    __syncwarp();
    if (act && mask4) {                                   // inverse columns, (x + 32) >> 6
#pragma unroll
        for (int j = 0; j < 2; j++) {
            int a[8], o[8];
            const int x = qx + 2 * r + j;
#pragma unroll
            for (int y = 0; y < 8; y++) a[y] = sm[(qy + y) * 16 + x];
            idct8_1d(a, o);
#pragma unroll
            for (int y = 0; y < 8; y++) sm[(qy + y) * 16 + x] = (o[y] + 32) >> 6;
        }
    }
    __syncwarp();
    if (act) {
        if (mask4) {
            int res[16];
#pragma unroll
            for (int i = 0; i < 16; i++) res[i] = sm[(by + (i >> 2)) * 16 + bx + (i & 3)];
            store_rec4x4(rec, rpitch, pred, res);
        } else {
            store_rec4x4(rec, rpitch, pred, nullptr);
        }
    }
    __syncwarp();
    return act ? (mask4 >> r) & 1 : 0;
}

template <bool T8>
__global__ void __launch_bounds__(K5_WARPS * 32)
k5_decide_inter_kernel(FramePlanes fp, int mbw, int mbh, int nmb_total, int is_p, int do_intra, int qp,
                       const b2_mv_t *__restrict__ mvq, const uint32_t *__restrict__ cost_inter,
                       const uint32_t *__restrict__ c16, const uint32_t *__restrict__ c4, const uint32_t *__restrict__ c8,
                       b2_mbinfo_t *__restrict__ info, b2_mbcoef_t *__restrict__ coef, b2_mv_t *__restrict__ prev_mv_out,
                       const uint8_t *__restrict__ pred_y, const uint8_t *__restrict__ part, const b2_mv_t *__restrict__ mv8)
{
    const int lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * K5_WARPS + (threadIdx.x >> 5);
    if (mbi >= nmb_total) return;
    const int per_frame = mbw * mbh;
    const int frame = mbi / per_frame, r = mbi - frame * per_frame;
    const int mby = r / mbw, mbx = r - mby * mbw;

    // K4: decision (identical on all lanes)
    uint32_t ci = 0xffffffffu; int it = B2_MB_I16x16;
    if (do_intra) {
        ci = c16[mbi];
        const uint32_t c = c4[mbi];
        if (c < ci) { ci = c; it = B2_MB_I4x4; }
        if (T8) { const uint32_t c8v = c8[mbi]; if (c8v < ci) { ci = c8v; it = B2_MB_I8x8; } }
    }
    const uint32_t cinter = is_p ? cost_inter[mbi] : 0u;
    const bool inter = is_p && !(do_intra && ci < cinter);
    b2_mv_t mv = {0, 0};
    if (inter) mv = mvq[mbi];
    // inter partitions (row N1): K2 chose the shape and one vector per 8x8 quadrant; quadrant 0 is `mv`
    const int shape = (inter && part != nullptr) ? part[mbi] : B2_PART_16x16;
    if (lane == 0) {
        if (shape != B2_PART_16x16) {
            info[mbi].part = (uint8_t)shape;
            info[mbi].mv8[0] = mv8[(size_t)mbi * 3 + 0]; info[mbi].mv8[1] = mv8[(size_t)mbi * 3 + 1]; info[mbi].mv8[2] = mv8[(size_t)mbi * 3 + 2];
        }
        info[mbi].mvx = mv.x; info[mbi].mvy = mv.y;
        info[mbi].mb_type = (uint8_t)(inter ? B2_MB_P16x16 : it);
        info[mbi].cost = inter ? cinter : ci;
        prev_mv_out[mbi] = mv;
    }
    if (!inter) return;                                   // intra MBs are reconstructed by K7

    b2_mbcoef_t *cf = &coef[mbi];
    int flags = 0;
    int src[16], pred[16];
    const int bx = blk_x(lane & 15) * 4, by = blk_y(lane & 15) * 4;
    const size_t offy = (size_t)(B2_PAD + mby * 16 + by) * fp.pitch + B2_PAD + mbx * 16 + bx;
    if (lane < 16) {
        load_src4x4(fp.cur[0] + frame * fp.stride_y + offy, fp.pitch, src);
        // motion-compensated prediction of the chosen MV, produced by K2 from its half-sample planes
        load_src4x4(pred_y + (size_t)mbi * 256 + by * 16 + bx, 16, pred);
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) src[i] = pred[i] = 0;
    }
    bool use8 = false;
    if (T8) {
        // transform size like x264's non-RD analysis: 8x8 iff SA8D(16x16) < SATD(16x16) of the prediction error.
        // The 8x8 Hadamard of a quadrant = butterflies over the 4x4 Hadamards of its four sub-blocks (lanes 4q..4q+3).
        int d[16], t[16];
#pragma unroll
        for (int i = 0; i < 16; i++) d[i] = src[i] - pred[i];
        hadamard4x4(d, t);
        uint32_t s4 = 0, s8 = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s4 += abs(t[i]);
        s4 >>= 1;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            int o = __shfl_xor_sync(0xffffffffu, t[i], 1);
            int u = (lane & 1) ? o - t[i] : t[i] + o;
            o = __shfl_xor_sync(0xffffffffu, u, 2);
            u = (lane & 2) ? o - u : u + o;
            s8 += abs(u);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s4 += __shfl_xor_sync(0xffffffffu, s4, o); s8 += __shfl_xor_sync(0xffffffffu, s8, o); }
        use8 = ((s8 + 2) >> 2) < s4;
    }
    // chroma blocks on lanes 16-23: source and motion-compensated prediction replace the zeros in src / pred
    const bool actc = lane >= 16 && lane < 24;
    const int pl = (lane >> 2) & 1, kc = lane & 3;
    const int qpc = chroma_qp(qp);
    const int cbx = (kc & 1) * 4, cby = (kc >> 1) * 4;
    const size_t offc = (size_t)(B2_PADC + mby * 8 + cby) * fp.pitchc + B2_PADC + mbx * 8 + cbx;
    if (actc) {
        load_src4x4((pl ? fp.cur[2] : fp.cur[1]) + frame * fp.stride_c + offc, fp.pitchc, src);   // selects: no local copy of fp
        // chroma 4x4 block k lies under luma quadrant k and moves with that quadrant's vector
        const b2_mv_t cmv = (shape != B2_PART_16x16 && kc > 0) ? mv8[(size_t)mbi * 3 + kc - 1] : mv;
        const uint8_t *rp = (pl ? fp.ref[2] : fp.ref[1]) + frame * fp.stride_c + offc + (ptrdiff_t)(cmv.y >> 3) * fp.pitchc + (cmv.x >> 3);
        const int fx = cmv.x & 7, fy = cmv.y & 7;
#pragma unroll
        for (int y = 0; y < 4; y++)
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const uint8_t *qq = rp + (size_t)y * fp.pitchc + x;
                pred[y * 4 + x] = ((8 - fx) * (8 - fy) * qq[0] + fx * (8 - fy) * qq[1] + (8 - fx) * fy * qq[fp.pitchc] +
                                   fx * fy * qq[fp.pitchc + 1] + 32) >> 6;
            }
    }
    uint8_t *recc = (pl ? fp.rec[2] : fp.rec[1]) + frame * fp.stride_c + offc;
    if (T8 && use8) {
        __shared__ int s_t8[K5_WARPS][256];
        flags = code_luma8x8_quad(lane, s_t8[threadIdx.x >> 5], src, pred, qp, cf->blk[lane & 15], fp.rec[0] + frame * fp.stride_y + offy,
                                  fp.pitch);
        const QParams q = make_qparams(qpc, false);
        const int cfl = code_chroma4x4(lane, actc, src, pred, q, qpc, cf, recc, fp.pitchc);
        if (actc) flags = cfl;
    } else {
        // 4x4 transform: luma (lanes 0-15) and chroma (lanes 16-23) blocks share one transform chain
        const bool is_c = lane >= 16;
        const QParams q = make_qparams(is_c ? qpc : qp, false);
        flags = code_mb4x4_mixed(lane, lane < 24, is_c, src, pred, q, qpc, is_c ? cf->blk[16 + 4 * pl + kc] : cf->blk[lane], cf,
                                 is_c ? recc : fp.rec[0] + frame * fp.stride_y + offy, is_c ? fp.pitchc : fp.pitch);
        if (lane >= 24) flags = 0;
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
        // transform_size_8x8_flag is only sent with coded luma (7.3.5): inferred 0 otherwise
        if (T8) info[mbi].transform8x8 = (uint8_t)(use8 && (mask & 0xffffu));
    }
}

}  // namespace

int b2_launch_decide_inter(const uint8_t *const cur[3], const uint8_t *const ref[3], uint8_t *const rec[3], int pitch,
                           int pitchc, size_t stride_y, size_t stride_c, int mbw, int mbh, int nframes, int is_p,
                           int do_intra, int qp, const b2_mv_t *d_mvq, const uint32_t *d_cost_inter, const uint32_t *d_c16,
                           const uint32_t *d_c4, const uint32_t *d_c8, b2_mbinfo_t *d_info, b2_mbcoef_t *d_coef, b2_mv_t *d_prev_mv,
                           const uint8_t *d_pred_y, int transform8x8, const uint8_t *d_part, const b2_mv_t *d_mv8, cudaStream_t st)
{
    FramePlanes fp;
    for (int i = 0; i < 3; i++) { fp.cur[i] = cur[i]; fp.ref[i] = ref ? ref[i] : nullptr; fp.rec[i] = rec[i]; }
    fp.pitch = pitch; fp.pitchc = pitchc; fp.stride_y = stride_y; fp.stride_c = stride_c;
    const int nmb = mbw * mbh * nframes;
    if (transform8x8)
        k5_decide_inter_kernel<true><<<(nmb + K5_WARPS - 1) / K5_WARPS, K5_WARPS * 32, 0, st>>>(
            fp, mbw, mbh, nmb, is_p, do_intra, qp, d_mvq, d_cost_inter, d_c16, d_c4, d_c8, d_info, d_coef, d_prev_mv, d_pred_y, d_part, d_mv8);
    else
        k5_decide_inter_kernel<false><<<(nmb + K5_WARPS - 1) / K5_WARPS, K5_WARPS * 32, 0, st>>>(
            fp, mbw, mbh, nmb, is_p, do_intra, qp, d_mvq, d_cost_inter, d_c16, d_c4, nullptr, d_info, d_coef, d_prev_mv, d_pred_y, d_part, d_mv8);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

