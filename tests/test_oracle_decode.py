"""Pins the NORMATIVE half of the oracle (intra predictors, luma/chroma interpolation, dequant,
IDCT, reconstruction) and the host CAVLC writer: the oracle's bitstream is decoded by libavcodec's
native H.264 decoder (bundled with the OpenCV wheel) and every decoded plane must equal the
oracle's reconstruction bit for bit (T3 'decoder drift test' of SURVEY.md 4).  CPU only."""
import collections
import numpy as np
import pytest


def smooth_seq(w, h, n, seed=0, cut=None):
    """smooth content moving by (5,3) quarter-pels per frame, optional scene cut -> exercises
    sub-pel MC, chroma MC and intra MBs inside P frames."""
    import cv2
    rng = np.random.default_rng(seed)

    def base():
        b = rng.integers(0, 256, ((h + 64) // 4 + 8, (w + 64) // 4 + 8)).astype(np.float32)
        return cv2.resize(b, ((w + 64) * 4, (h + 64) * 4), interpolation=cv2.INTER_CUBIC)
    B, B2 = base(), base()
    frames = []
    for t in range(n):
        src = B2 if (cut is not None and t >= cut) else B
        ox, oy = 5 * t + 3, 3 * t + 1
        img = src[oy:oy + 4 * h:4, ox:ox + 4 * w:4]
        img = np.clip(img + rng.normal(0, 1.5, img.shape), 0, 255).astype(np.uint8)
        u = (img[::2, ::2] // 2 + 64).astype(np.uint8)
        v = (255 - img[::2, ::2] // 2 - 30).astype(np.uint8)
        frames.append((img, u, v))
    return frames


def coarse_seq(w, h, n, seed=0, scale=10, noise=1.0):
    """large smooth structures (random field upsampled by `scale`) + a little noise, panning by (3,2) px per frame:
    the content on which intra 8x8 prediction and the 8x8 transform win"""
    import cv2
    rng = np.random.default_rng(seed)
    b = rng.integers(0, 256, ((h + 64) // scale + 4, (w + 64) // scale + 4)).astype(np.float32)
    B = cv2.resize(b, (b.shape[1] * scale, b.shape[0] * scale), interpolation=cv2.INTER_CUBIC)
    frames = []
    for t in range(n):
        ox, oy = 3 * t + 2, 2 * t + 1
        img = np.clip(B[oy:oy + h, ox:ox + w] + rng.normal(0, noise, (h, w)), 0, 255).astype(np.uint8)
        u = (img[::2, ::2] // 2 + 64).astype(np.uint8)[:(h + 1) // 2, :(w + 1) // 2]
        v = (255 - img[::2, ::2] // 2 - 30).astype(np.uint8)[:(h + 1) // 2, :(w + 1) // 2]
        frames.append((img, u, v))
    return frames


def shear_seq(w, h, n, seed=0, stripe=24, band=24, amp=1):
    """content whose vertical stripes / horizontal bands move by different quarter-pel amounts per frame: macroblocks
    that straddle a boundary are better predicted with 8x16 / 16x8 / 8x8 partitions (row N1)"""
    import cv2
    rng = np.random.default_rng(seed)
    b = rng.integers(0, 256, ((h + 64) // 5 + 8, (w + 64) // 5 + 8)).astype(np.float32)
    B = cv2.resize(b, ((w + 64) * 4, (h + 64) * 4), interpolation=cv2.INTER_CUBIC)
    X = np.arange(w); Y = np.arange(h)
    sx = (X // stripe) % 3 - 1; sy = (Y // band) % 3 - 1              # -1, 0, +1 quarter-pels per frame
    frames = []
    for t in range(n):
        xi = 4 * X + 64 + t * 2 * amp * sx + 3 * t                     # common pan + per-stripe shear (amp 1: +-1/2 pel per frame)
        yi = 4 * Y + 64 + t * 2 * amp * sy + 2 * t
        img = np.clip(B[np.ix_(yi, xi)] + rng.normal(0, 0.8, (h, w)), 0, 255).astype(np.uint8)
        u = (img[::2, ::2] // 2 + 64).astype(np.uint8)[:(h + 1) // 2, :(w + 1) // 2]
        v = (255 - img[::2, ::2] // 2 - 30).astype(np.uint8)[:(h + 1) // 2, :(w + 1) // 2]
        frames.append((img, u, v))
    return frames


def _roundtrip(oracle, frames, w, h, **kw):
    bs, recons, infos, coefs = oracle.encode_sequence(frames, w, h, **kw)
    dec = oracle.decode_yuv(oracle.split_access_units(bs))
    assert len(dec) == len(frames)
    ch, cw = (h + 1) // 2, (w + 1) // 2
    for i, (dy, du, dv) in enumerate(dec):
        r = recons[i]
        assert np.array_equal(dy, r.y[:h, :w]), f"luma drift in frame {i}"
        assert np.array_equal(du, r.u[:ch, :cw]), f"U drift in frame {i}"
        assert np.array_equal(dv, r.v[:ch, :cw]), f"V drift in frame {i}"
    return bs, recons, infos


@pytest.mark.parametrize("w,h,qp,R,cut", [(176, 144, 26, 16, None), (320, 240, 32, 32, 4), (208, 160, 18, 16, 3),
                                          (176, 144, 40, 16, 2), (318, 242, 28, 16, None), (64, 48, 12, 16, 1),
                                          (96, 80, 51, 16, 2)])
def test_decoder_matches_oracle_recon(oracle, w, h, qp, R, cut):
    frames = smooth_seq(w, h, 6, seed=qp, cut=cut)
    bs, recons, infos = _roundtrip(oracle, frames, w, h, qp=qp, merange=R, gop=32)
    types = collections.Counter()
    for inf in infos[1:]:
        types.update(inf["mb_type"].tolist())
    assert types[0] > 0                                   # inter MBs present
    if cut is not None:
        assert types[1] + types[2] > 0                    # the scene cut produced intra MBs in a P frame


def test_synthetic_pan_and_gops(oracle):
    """the bench content (integer pan): two closed GOPs, skip MBs, I16x16 and I4x4 MBs"""
    w, h = 320, 240
    frames = [oracle.synth_frame(w, h, t) for t in range(8)]
    bs, recons, infos = _roundtrip(oracle, frames, w, h, qp=30, merange=16, gop=4)
    assert set(np.unique(infos[0]["mb_type"])) <= {1, 2}
    mv = infos[1]
    interior = (mv["mvx"].reshape(15, 20)[:-1, :-1] == 12).mean()
    assert interior > 0.9                                 # known-answer MV (+3,+2) px = (12,8) qpel


def test_flat_and_extreme_content(oracle):
    """flat frames (all skip / DC-only), saturated noise (large levels, escape codes)"""
    w, h = 64, 64
    rng = np.random.default_rng(3)
    flat = [(np.full((h, w), 90, np.uint8), np.full((h // 2, w // 2), 128, np.uint8), np.full((h // 2, w // 2), 128, np.uint8))] * 3
    _roundtrip(oracle, flat, w, h, qp=26, merange=16, gop=8)
    noise = [((rng.integers(0, 2, (h, w)) * 255).astype(np.uint8), (rng.integers(0, 2, (h // 2, w // 2)) * 255).astype(np.uint8),
              (rng.integers(0, 2, (h // 2, w // 2)) * 255).astype(np.uint8)) for _ in range(3)]
    for qp in (10, 20, 45):
        _roundtrip(oracle, noise, w, h, qp=qp, merange=16, gop=8)


def test_psnr_monotone_in_qp(oracle):
    w, h = 176, 144
    frames = smooth_seq(w, h, 4, seed=5)
    last_psnr, last_size = 100.0, 1 << 30
    for qp in (20, 28, 36, 44):
        bs, recons, _ = _roundtrip(oracle, frames, w, h, qp=qp, merange=16, gop=32)
        src = oracle.OFrame(w, h).load(*frames[-1])
        psnr = oracle.lib().b2o_psnr_y(oracle.C.byref(src.f), oracle.C.byref(recons[-1].f))
        assert psnr < last_psnr and len(bs) < last_size
        last_psnr, last_size = psnr, len(bs)


@pytest.mark.parametrize("w,h,qp,R,cut", [(176, 144, 26, 16, None), (320, 240, 36, 32, 3), (208, 160, 18, 16, 2), (96, 80, 51, 16, 2),
                                          (318, 242, 30, 16, None), (64, 48, 12, 16, 1)])
def test_decoder_matches_oracle_recon_with_deblocking(oracle, w, h, qp, R, cut):
    """in-loop deblocking filter (8.7) of the oracle against libavcodec's"""
    frames = smooth_seq(w, h, 6, seed=qp, cut=cut)
    _roundtrip(oracle, frames, w, h, qp=qp, merange=R, gop=32, deblock=1)


@pytest.mark.parametrize("offs", [(-1, -1), (1, 1), (-3, -3), (2, -2), (-6, 6), (6, -6)])
@pytest.mark.parametrize("cabac", [0, 1])
def test_deblock_offsets_decoder_exact(oracle, offs, cabac):
    """slice_alpha_c0_offset_div2 / slice_beta_offset_div2 (x264's tunes: film -1:-1 -- the reference's default, av_encode.c:103 --
    animation 1:1, stillimage -3:-3): the oracle's loop filter with offsets == libavcodec's, and the offsets change the picture"""
    w, h, qp = 176, 144, 33
    frames = smooth_seq(w, h, 5, seed=17, cut=3)
    bs, recons, _ = _roundtrip(oracle, frames, w, h, qp=qp, merange=16, gop=32, deblock=1, cabac=cabac, deblock_offsets=offs)
    bs0, recons0, _ = _roundtrip(oracle, frames, w, h, qp=qp, merange=16, gop=32, deblock=1, cabac=cabac)
    assert any(not np.array_equal(a.y, b.y) for a, b in zip(recons, recons0))
