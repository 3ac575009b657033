// b2_mbcode.cuh -- warp-level macroblock coding helpers shared by K5 (inter) and K7 (intra):
// residual -> DCT -> quant -> levels -> dequant -> IDCT -> reconstruction for 4x4 blocks owned by lanes.
#pragma once
#include "b2_h264.cuh"

namespace b2 {

struct FramePlanes {
    const uint8_t *cur[3];
    const uint8_t *ref[3];
    uint8_t *rec[3];
    int pitch, pitchc;
    size_t stride_y, stride_c;
};

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
        w[0] = ((fi * q.ls[0]) * (1 << (qpc / 6))) >> 5;
        idct4x4(w);
        store_rec4x4(rec, rpitch, pred, w);
    }
    return (nnz ? 1 : 0) | (zdc ? 2 : 0);
}

// One transform chain for a whole inter macroblock with the 4x4 transform: lanes 0-15 carry the luma blocks, lanes 16-23 the
// chroma blocks (pl = (lane-16)>>2, k = lane&3) -- same arithmetic as code_luma4x4 on the former and code_chroma4x4 on the
// latter (the DC coefficient of a chroma lane leaves the AC quantiser and goes through the 2x2 Hadamard by shuffle), but
// the warp walks DCT / quantiser / scaling / inverse once instead of once for luma and once more for chroma.
// q = this lane's quantiser (luma QP or chroma QP).  ALL 32 lanes must call; `act` lanes touch memory.  Returns bit0 = AC
// (luma: any level) non-zero, bit1 = chroma DC non-zero.
__device__ __forceinline__ int code_mb4x4_mixed(int lane, bool act, bool is_chroma, const int src[16], const int pred[16],
                                                const QParams &q, int qpc, int16_t *lev, b2_mbcoef_t *coef, uint8_t *rec, int rpitch)
{
    const int pl = (lane >> 2) & 1, k = lane & 3, base = lane & ~3;
    int w[16], z[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = src[i] - pred[i];
    dct4x4(w);
    const int v0 = __shfl_sync(0xffffffffu, w[0], base + 0), v1 = __shfl_sync(0xffffffffu, w[0], base + 1);
    const int v2 = __shfl_sync(0xffffffffu, w[0], base + 2), v3 = __shfl_sync(0xffffffffu, w[0], base + 3);
    const int f = k == 0 ? v0 + v1 + v2 + v3 : k == 1 ? v0 - v1 + v2 - v3 : k == 2 ? v0 + v1 - v2 - v3 : v0 - v1 - v2 + v3;
    const int zdc = is_chroma ? quant_dc(f, q) : 0;
    int nnz = quant4x4(w, z, q, false);
    if (is_chroma) { nnz -= z[0] != 0; z[0] = 0; }
    const int g0 = __shfl_sync(0xffffffffu, zdc, base + 0), g1 = __shfl_sync(0xffffffffu, zdc, base + 1);
    const int g2 = __shfl_sync(0xffffffffu, zdc, base + 2), g3 = __shfl_sync(0xffffffffu, zdc, base + 3);
    const int fi = k == 0 ? g0 + g1 + g2 + g3 : k == 1 ? g0 - g1 + g2 - g3 : k == 2 ? g0 + g1 - g2 - g3 : g0 - g1 - g2 + g3;
    if (act) {
        store_levels_zigzag(lev, z);
        if (is_chroma) coef->blk[25][4 * pl + k] = (int16_t)zdc;
        dequant4x4(z, w, q, false);                      // all-zero levels scale and invert to a zero residual
        if (is_chroma) w[0] = ((fi * q.ls[0]) * (1 << (qpc / 6))) >> 5;
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

}  // namespace b2
