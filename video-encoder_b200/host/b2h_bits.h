/*
 * b2h_bits.h -- bit writer and NAL packing shared by the host entropy writers (b2h_cavlc.c, b2h_cabac.c).
 * Internal to the host stage; not part of the C-ABI.  ITU-T H.264 7.2 (bit order), 7.4.1 (emulation prevention).
 */
#ifndef B2H_BITS_H
#define B2H_BITS_H
#include <stddef.h>
#include <stdint.h>

/* ---- bit writer ----------------------------------------------------------------------------*/
typedef struct {
    uint8_t *buf;
    size_t cap, pos;
    uint64_t acc;
    int nbits;
    int overflow;
} bs_t;

static inline void bs_init(bs_t *b, uint8_t *buf, size_t cap) { b->buf = buf; b->cap = cap; b->pos = 0; b->acc = 0; b->nbits = 0; b->overflow = 0; }

static inline void bs_put(bs_t *b, int n, uint32_t v)
{
    b->acc = (b->acc << n) | (v & (n == 32 ? 0xffffffffu : ((1u << n) - 1)));
    b->nbits += n;
    while (b->nbits >= 8) {
        b->nbits -= 8;
        if (b->pos < b->cap) b->buf[b->pos++] = (uint8_t)(b->acc >> b->nbits);
        else b->overflow = 1;
    }
}
static inline void bs_ue(bs_t *b, uint32_t v)
{
    uint32_t x = v + 1;
    int len = 0;
    while ((x >> len) > 1) len++;
    if (len) bs_put(b, len, 0);
    bs_put(b, len + 1, x);
}
static inline void bs_se(bs_t *b, int v) { bs_ue(b, v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }
static inline void bs_trailing(bs_t *b)
{
    bs_put(b, 1, 1);
    if (b->nbits) bs_put(b, 8 - b->nbits, 0);
}

/* NAL = header + RBSP with emulation prevention (00 00 0x -> 00 00 03 0x) */
static inline size_t nal_pack(int ref_idc, int type, const uint8_t *rbsp, size_t n, uint8_t *out, size_t cap)
{
    size_t o = 0;
    int zeros = 0;
    if (cap < 1) return 0;
    out[o++] = (uint8_t)((ref_idc << 5) | type);
    for (size_t i = 0; i < n; i++) {
        if (zeros >= 2 && rbsp[i] <= 3) {
            if (o >= cap) return 0;
            out[o++] = 3;
            zeros = 0;
        }
        if (o >= cap) return 0;
        out[o++] = rbsp[i];
        zeros = rbsp[i] == 0 ? zeros + 1 : 0;
    }
    return o;
}

#endif
