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
#include "b2_h264.cuh"

namespace {

using namespace b2;

__device__ __forceinline__ void load_src4x4(const uint8_t *p, int pitch, int s[16])
{
#pragma unroll
    for (int y = 0; y < 4; y++) {
        uint32_t w = *(const uint32_t *)(p + (size_t)y * pitch);
#pragma unroll
        for (int x = 0; x < 4; x++) s[y * 4 + x] = (w >> (8 * x)) & 255;
    }
}
__device__ __forceinline__ void store_rec4x4(uint8_t *p, int pitch, const int pred[16], const int *res)
{
#pragma unroll
    for (int y = 0; y < 4; y++) {
        uint32_t w = 0;
#pragma unroll
        for (int x = 0; x < 4; x++) {
            int v = pred[y * 4 + x] + (res ? res[y * 4 + x] : 0);
            w |= (uint32_t)b2_clip255(v) << (8 * x);
        }
        *(uint32_t *)(p + (size_t)y * pitch) = w;
    }
}

// plain luma 4x4 (inter or I4x4): returns true when the block has a non-zero level
__device__ __forceinline__ bool code_luma4x4(const int src[16], const int pred[16], const QParams &q, int16_t *lev,
                                             uint8_t *rec, int rpitch)
{
    int w[16], z[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = src[i] - pred[i];
    dct4x4(w);
    const int nnz = quant4x4(w, z, q, false);
    store_levels_zigzag(lev, z);
    if (nnz) {
        dequant4x4(z, w, q, false);
        idct4x4(w);
        store_rec4x4(rec, rpitch, pred, w);
    } else {
        store_rec4x4(rec, rpitch, pred, nullptr);
    }
    return nnz != 0;
}

// chroma 4x4 blocks on lanes 16..23 (pl = (lane-16)>>2, k = lane&3).  ALL 32 lanes must call
// (shuffles); only `act` lanes touch memory.  Returns bit0 = AC non-zero, bit1 = DC non-zero.
__device__ __forceinline__ int code_chroma4x4(int lane, bool act, const int src[16], const int pred[16],
                                              const QParams &q, int qpc, b2_mbcoef_t *coef, uint8_t *rec, int rpitch)
{
    const int pl = (lane >> 2) & 1, k = lane & 3, base = lane & ~3;
    int w[16], z[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = src[i] - pred[i];
    dct4x4(w);
    const int v0 = __shfl_sync(0xffffffffu, w[0], base + 0), v1 = __shfl_sync(0xffffffffu, w[0], base + 1);
    const int v2 = __shfl_sync(0xffffffffu, w[0], base + 2), v3 = __shfl_sync(0xffffffffu, w[0], base + 3);
    const int f = k == 0 ? v0 + v1 + v2 + v3 : k == 1 ? v0 - v1 + v2 - v3 : k == 2 ? v0 + v1 - v2 - v3 : v0 - v1 - v2 + v3;
    const int zdc = quant_dc(f, q);
    const int nnz = quant4x4(w, z, q, true);
    const int g0 = __shfl_sync(0xffffffffu, zdc, base + 0), g1 = __shfl_sync(0xffffffffu, zdc, base + 1);
    const int g2 = __shfl_sync(0xffffffffu, zdc, base + 2), g3 = __shfl_sync(0xffffffffu, zdc, base + 3);
    const int fi = k == 0 ? g0 + g1 + g2 + g3 : k == 1 ? g0 - g1 + g2 - g3 : k == 2 ? g0 + g1 - g2 - g3 : g0 - g1 - g2 + g3;
    if (act) {
        store_levels_zigzag(coef->blk[16 + 4 * pl + k], z);
        coef->blk[25][4 * pl + k] = (int16_t)zdc;
        dequant4x4(z, w, q, true);
        w[0] = ((fi * q.ls[0]) << (qpc / 6)) >> 5;
        idct4x4(w);
        store_rec4x4(rec, rpitch, pred, w);
    }
    return (nnz ? 1 : 0) | (zdc ? 2 : 0);
}

__device__ __forceinline__ int cbp_from_mask(int mb_type, uint32_t mask)
{
    int cbp = 0;
    if (mb_type == B2_MB_I16x16) {
        if (mask & 0xffffu) cbp = 15;
    } else {
#pragma unroll
        for (int qd = 0; qd < 4; qd++)
            if (mask & (0xfu << (4 * qd))) cbp |= 1 << qd;
    }
    if (mask & 0x00ff0000u) cbp |= 2 << 4;
    else if (mask & 0x06000000u) cbp |= 1 << 4;
    return cbp;
}

// nnz mask from per-lane flags: lanes 0-15 luma (flag bit0), lanes 16-23 chroma (bit0 AC, bit1 DC)
__device__ __forceinline__ uint32_t gather_mask(int lane, int flags, bool luma_dc)
{
    const uint32_t ac = __ballot_sync(0xffffffffu, flags & 1);
    const uint32_t dc = __ballot_sync(0xffffffffu, (flags & 2) && lane >= 16 && lane < 24);
    uint32_t mask = ac & 0x00ffffffu;
    if (dc & 0x000f0000u) mask |= 1u << 25;
    if (dc & 0x00f00000u) mask |= 1u << 26;
    if (luma_dc) mask |= 1u << 24;
    return mask;
}

// ---- K5: decision + inter reconstruction ---------------------------------------------------------
struct FramePlanes {
    const uint8_t *cur[3];
    const uint8_t *ref[3];
    uint8_t *rec[3];
    int pitch, pitchc;
    size_t stride_y, stride_c;
};

constexpr int K5_WARPS = 4;

__global__ void __launch_bounds__(K5_WARPS * 32)
k5_decide_inter_kernel(FramePlanes fp, int mbw, int mbh, int nmb_total, int is_p, int do_intra, int qp,
                       const b2_mv_t *__restrict__ mvq, const uint32_t *__restrict__ cost_inter,
                       const uint32_t *__restrict__ c16, const uint32_t *__restrict__ c4,
                       b2_mbinfo_t *__restrict__ info, b2_mbcoef_t *__restrict__ coef, b2_mv_t *__restrict__ prev_mv_out)
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
        const uint8_t *rp = fp.ref[0] + frame * fp.stride_y + off + (ptrdiff_t)(mv.y >> 2) * fp.pitch + (mv.x >> 2);
        const int fx = mv.x & 3, fy = mv.y & 3;
#pragma unroll
        for (int y = 0; y < 4; y++)
#pragma unroll
            for (int x = 0; x < 4; x++) pred[y * 4 + x] = qpel_sample(rp + (size_t)y * fp.pitch + x, fp.pitch, fx, fy);
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

// ---- K7: intra reconstruction, wavefront over anti-diagonals, one CTA per frame ----------------------
constexpr int K7_WARPS = 16;

// z-order index of the block at (x,y) in units of 4 px
__device__ __forceinline__ int zidx(int x, int y) { return (x & 1) | ((y & 1) << 1) | ((x >> 1) << 2) | ((y >> 1) << 3); }

__device__ void k7_code_intra_mb(int lane, const FramePlanes &fp, int frame, int mbx, int mby, int mbw, int qp,
                                 b2_mbinfo_t *mi, b2_mbcoef_t *cf)
{
    const int mba = mb_avail(mbx, mby, mbw);
    const bool hasT = mba & 2, hasL = mba & 1;
    const int mb_type = mi->mb_type;
    const size_t offy = (size_t)(B2_PAD + mby * 16) * fp.pitch + B2_PAD + mbx * 16;
    const uint8_t *sy = fp.cur[0] + frame * fp.stride_y + offy;
    uint8_t *ry = fp.rec[0] + frame * fp.stride_y + offy;
    int flags = 0;
    bool luma_dc = false;

    // ---- chroma first (independent of the luma path) ----
    {
        const bool act = lane >= 16 && lane < 24;
        const int pl = (lane >> 2) & 1, k = lane & 3, cbx = (k & 1) * 4, cby = (k >> 1) * 4;
        const int qpc = chroma_qp(qp);
        const QParams q = make_qparams(qpc, true);
        const size_t offc = (size_t)(B2_PADC + mby * 8) * fp.pitchc + B2_PADC + mbx * 8;
        uint8_t *rc = fp.rec[1 + pl] + frame * fp.stride_c + offc;
        int src[16], pred[16];
#pragma unroll
        for (int i = 0; i < 16; i++) src[i] = pred[i] = 0;
        if (act) {
            const int mode = mi->chroma_mode;
            load_src4x4(fp.cur[1 + pl] + frame * fp.stride_c + offc + (size_t)cby * fp.pitchc + cbx, fp.pitchc, src);
            int top[8], left[8], tl = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) { top[i] = hasT ? rc[-(ptrdiff_t)fp.pitchc + i] : 0; left[i] = hasL ? rc[(size_t)i * fp.pitchc - 1] : 0; }
            if (mba & 4) tl = rc[-(ptrdiff_t)fp.pitchc - 1];
            if (mode == B2_IC_DC) {
                const int t0 = top[0] + top[1] + top[2] + top[3], t1 = top[4] + top[5] + top[6] + top[7];
                const int l0 = left[0] + left[1] + left[2] + left[3], l1 = left[4] + left[5] + left[6] + left[7];
                int dcv;
                if (k == 0) dcv = (hasT && hasL) ? (t0 + l0 + 4) >> 3 : hasT ? (t0 + 2) >> 2 : hasL ? (l0 + 2) >> 2 : 128;
                else if (k == 1) dcv = hasT ? (t1 + 2) >> 2 : hasL ? (l0 + 2) >> 2 : 128;
                else if (k == 2) dcv = hasL ? (l1 + 2) >> 2 : hasT ? (t0 + 2) >> 2 : 128;
                else dcv = (hasT && hasL) ? (t1 + l1 + 4) >> 3 : hasT ? (t1 + 2) >> 2 : hasL ? (l1 + 2) >> 2 : 128;
#pragma unroll
                for (int i = 0; i < 16; i++) pred[i] = dcv;
            } else if (mode == B2_IC_H) {
#pragma unroll
                for (int i = 0; i < 16; i++) pred[i] = left[cby + (i >> 2)];
            } else if (mode == B2_IC_V) {
#pragma unroll
                for (int i = 0; i < 16; i++) pred[i] = top[cbx + (i & 3)];
            } else {
                int Hc = 0, Vc = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    Hc += (i + 1) * (top[4 + i] - (i == 3 ? tl : top[2 - i]));
                    Vc += (i + 1) * (left[4 + i] - (i == 3 ? tl : left[2 - i]));
                }
                const int pa = 16 * (left[7] + top[7]), pb = (34 * Hc + 32) >> 6, pc = (34 * Vc + 32) >> 6;
#pragma unroll
                for (int i = 0; i < 16; i++)
                    pred[i] = b2_clip255((pa + pb * (cbx + (i & 3) - 3) + pc * (cby + (i >> 2) - 3) + 16) >> 5);
            }
        }
        // all chroma lanes must have read their neighbours before anyone overwrites the MB (they only
        // read outside the MB, so no hazard inside the warp)
        const int cfl = code_chroma4x4(lane, act, src, pred, q, qpc, cf, rc + (size_t)cby * fp.pitchc + cbx, fp.pitchc);
        if (act) flags = cfl;
    }

    const QParams q = make_qparams(qp, true);
    if (mb_type == B2_MB_I16x16) {
        const int mode = mi->i16_mode;
        const int l16 = lane & 15;
        const int topv = hasT ? ry[-(ptrdiff_t)fp.pitch + l16] : 0;
        const int leftv = hasL ? ry[(size_t)l16 * fp.pitch - 1] : 0;
        const int tlv = (mba & 4) ? ry[-(ptrdiff_t)fp.pitch - 1] : 0;
        const int bx = blk_x(l16) * 4, by = blk_y(l16) * 4;
        int t4[4], l4[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { t4[i] = __shfl_sync(0xffffffffu, topv, bx + i); l4[i] = __shfl_sync(0xffffffffu, leftv, by + i); }
        int sumT = lane < 16 ? topv : 0, sumL = lane < 16 ? leftv : 0;
        int hterm = 0, vterm = 0;
        {
            const int i = l16 & 7;
            const int ta = __shfl_sync(0xffffffffu, topv, 8 + i), tb = __shfl_sync(0xffffffffu, topv, (6 - i) & 15);
            const int la = __shfl_sync(0xffffffffu, leftv, 8 + i), lb = __shfl_sync(0xffffffffu, leftv, (6 - i) & 15);
            if (lane < 8) { hterm = (i + 1) * (ta - (i == 7 ? tlv : tb)); vterm = (i + 1) * (la - (i == 7 ? tlv : lb)); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sumT += __shfl_xor_sync(0xffffffffu, sumT, o); sumL += __shfl_xor_sync(0xffffffffu, sumL, o);
            hterm += __shfl_xor_sync(0xffffffffu, hterm, o); vterm += __shfl_xor_sync(0xffffffffu, vterm, o);
        }
        const int t15 = __shfl_sync(0xffffffffu, topv, 15), l15 = __shfl_sync(0xffffffffu, leftv, 15);
        const int pa = 16 * (l15 + t15), pb = (5 * hterm + 32) >> 6, pc = (5 * vterm + 32) >> 6;
        const int dc = (hasT && hasL) ? (sumT + sumL + 16) >> 5 : (hasT || hasL) ? (sumT + sumL + 8) >> 4 : 128;
        int src[16], pred[16], w[16], z[16];
#pragma unroll
        for (int i = 0; i < 16; i++) src[i] = pred[i] = 0;
        if (lane < 16) {
            load_src4x4(sy + (size_t)by * fp.pitch + bx, fp.pitch, src);
#pragma unroll
            for (int y = 0; y < 4; y++)
#pragma unroll
                for (int x = 0; x < 4; x++)
                    pred[y * 4 + x] = mode == B2_I16_V ? t4[x] : mode == B2_I16_H ? l4[y] : mode == B2_I16_DC ? dc
                                      : b2_clip255((pa + pb * (bx + x - 7) + pc * (by + y - 7) + 16) >> 5);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = src[i] - pred[i];
        dct4x4(w);
        // 4x4 Hadamard of the 16 DCs: lane r (< 16) produces output raster position r = v*4+u
        const int v = (l16 >> 2), u = l16 & 3;
        int acc = 0;
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const int dcs = __shfl_sync(0xffffffffu, w[0], zidx(s & 3, s >> 2));
            acc += h4_sign(v, s >> 2) * h4_sign(u, s & 3) * dcs;
        }
        const int zdc = quant_dc((acc + 1) >> 1, q);
        int acc2 = 0;
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const int zs = __shfl_sync(0xffffffffu, zdc, s);
            acc2 += h4_sign(v, s >> 2) * h4_sign(u, s & 3) * zs;
        }
        const int dq_r = q.s >= 6 ? (acc2 * q.ls[0]) << (q.s - 6) : (acc2 * q.ls[0] + (1 << (5 - q.s))) >> (6 - q.s);
        // block l16 sits at raster position (blk_y*4 + blk_x)
        const int my_dc = __shfl_sync(0xffffffffu, dq_r, blk_y(l16) * 4 + blk_x(l16));
        luma_dc = __ballot_sync(0xffffffffu, lane < 16 && zdc != 0) != 0;
        if (lane < 16) {
            // inverse zig-zag: raster r -> scan position
            const int izz[16] = {0, 1, 5, 6, 2, 4, 7, 12, 3, 8, 11, 13, 9, 10, 14, 15};
            cf->blk[24][izz[l16]] = (int16_t)zdc;
            const int nnz = quant4x4(w, z, q, true);
            store_levels_zigzag(cf->blk[l16], z);
            dequant4x4(z, w, q, true);
            w[0] = my_dc;
            idct4x4(w);
            store_rec4x4(ry + (size_t)by * fp.pitch + bx, fp.pitch, pred, w);
            flags = nnz ? 1 : 0;
        }
    } else {
        // I4x4: 10-step wavefront s = bx + 2*by over the sixteen blocks
        for (int s = 0; s < 10; s++) {
            if (lane < 16) {
                const int bx = blk_x(lane), by = blk_y(lane);
                if (bx + 2 * by == s) {
                    const int ba = blk_avail(lane, mba);
                    uint8_t *rb = ry + (size_t)(by * 4) * fp.pitch + bx * 4;
                    int E[13], src[16], pred[16];
                    load_edge4x4(rb, fp.pitch, ba, E);
                    pred4x4(mi->i4_mode[lane], E, ba, pred);
                    load_src4x4(sy + (size_t)(by * 4) * fp.pitch + bx * 4, fp.pitch, src);
                    flags = code_luma4x4(src, pred, q, cf->blk[lane], rb, fp.pitch) ? 1 : 0;
                }
            }
            __syncwarp();
        }
        if (lane == 24) {
            uint4 z4 = make_uint4(0, 0, 0, 0);
            ((uint4 *)cf->blk[24])[0] = z4; ((uint4 *)cf->blk[24])[1] = z4;
        }
    }
    if (lane == 25) ((uint4 *)cf->blk[25])[1] = make_uint4(0, 0, 0, 0);
    const uint32_t mask = gather_mask(lane, flags, luma_dc);
    if (lane == 0) {
        mi->cbp = (uint8_t)cbp_from_mask(mb_type, mask);
        mi->nnz_mask = mask;
    }
}

__global__ void __launch_bounds__(K7_WARPS * 32)
k7_intra_wavefront_kernel(FramePlanes fp, int mbw, int mbh, int qp, b2_mbinfo_t *__restrict__ info,
                          b2_mbcoef_t *__restrict__ coef)
{
    extern __shared__ int s_diag_cnt[];                 // intra MBs per anti-diagonal
    const int frame = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ndiag = mbw + 2 * (mbh - 1);
    b2_mbinfo_t *finfo = info + (size_t)frame * mbw * mbh;
    b2_mbcoef_t *fcoef = coef + (size_t)frame * mbw * mbh;
    for (int i = threadIdx.x; i < ndiag; i += blockDim.x) s_diag_cnt[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < mbw * mbh; i += blockDim.x)
        if (finfo[i].mb_type != B2_MB_P16x16) atomicAdd(&s_diag_cnt[(i % mbw) + 2 * (i / mbw)], 1);
    __syncthreads();
    for (int d = 0; d < ndiag; d++) {
        if (s_diag_cnt[d] == 0) continue;               // uniform over the CTA
        // MBs on this diagonal: mby in [max(0, ceil((d-mbw+1)/2)), min(mbh-1, d/2)], mbx = d - 2*mby
        const int y_lo = max(0, (d - mbw + 2) >> 1), y_hi = min(mbh - 1, d >> 1);
        for (int mby = y_lo + warp; mby <= y_hi; mby += K7_WARPS) {
            const int mbx = d - 2 * mby;
            const int i = mby * mbw + mbx;
            if (finfo[i].mb_type != B2_MB_P16x16)
                k7_code_intra_mb(lane, fp, frame, mbx, mby, mbw, qp, &finfo[i], &fcoef[i]);
        }
        __syncthreads();
    }
}

}  // namespace

int b2_launch_decide_inter(const uint8_t *const cur[3], const uint8_t *const ref[3], uint8_t *const rec[3], int pitch,
                           int pitchc, size_t stride_y, size_t stride_c, int mbw, int mbh, int nframes, int is_p,
                           int do_intra, int qp, const b2_mv_t *d_mvq, const uint32_t *d_cost_inter, const uint32_t *d_c16,
                           const uint32_t *d_c4, b2_mbinfo_t *d_info, b2_mbcoef_t *d_coef, b2_mv_t *d_prev_mv,
                           cudaStream_t st)
{
    FramePlanes fp;
    for (int i = 0; i < 3; i++) { fp.cur[i] = cur[i]; fp.ref[i] = ref ? ref[i] : nullptr; fp.rec[i] = rec[i]; }
    fp.pitch = pitch; fp.pitchc = pitchc; fp.stride_y = stride_y; fp.stride_c = stride_c;
    const int nmb = mbw * mbh * nframes;
    k5_decide_inter_kernel<<<(nmb + K5_WARPS - 1) / K5_WARPS, K5_WARPS * 32, 0, st>>>(
        fp, mbw, mbh, nmb, is_p, do_intra, qp, d_mvq, d_cost_inter, d_c16, d_c4, d_info, d_coef, d_prev_mv);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}

int b2_launch_intra_recon(const uint8_t *const cur[3], uint8_t *const rec[3], int pitch, int pitchc, size_t stride_y,
                          size_t stride_c, int mbw, int mbh, int nframes, int qp, b2_mbinfo_t *d_info,
                          b2_mbcoef_t *d_coef, cudaStream_t st)
{
    FramePlanes fp;
    for (int i = 0; i < 3; i++) { fp.cur[i] = cur[i]; fp.ref[i] = nullptr; fp.rec[i] = rec[i]; }
    fp.pitch = pitch; fp.pitchc = pitchc; fp.stride_y = stride_y; fp.stride_c = stride_c;
    const int ndiag = mbw + 2 * (mbh - 1);
    k7_intra_wavefront_kernel<<<nframes, K7_WARPS * 32, ndiag * sizeof(int), st>>>(fp, mbw, mbh, qp, d_info, d_coef);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
}
