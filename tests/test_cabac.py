"""Row N3: the host CABAC slice writer (video-encoder_b200/host/b2h_cabac.c).  The oracle's per-MB
decisions and levels are written as a Main-profile CABAC stream, decoded by libavcodec's native H.264 decoder and every
decoded plane must equal the oracle's reconstruction bit for bit: one wrong context index, binarisation or table entry
desynchronises the arithmetic decoder for the rest of the slice, so this pins the whole writer.  CPU only."""
import collections
import numpy as np
import pytest
from test_oracle_decode import smooth_seq, coarse_seq, _roundtrip


@pytest.mark.parametrize("w,h,qp,R,cut,deblock", [(176, 144, 26, 16, None, 0), (320, 240, 32, 32, 4, 1), (208, 160, 18, 16, 3, 0),
                                                  (176, 144, 40, 16, 2, 1), (318, 242, 28, 16, None, 0), (64, 48, 12, 16, 1, 0),
                                                  (96, 80, 51, 16, 2, 1)])
def test_cabac_stream_decodes_to_oracle_recon(oracle, w, h, qp, R, cut, deblock):
    frames = smooth_seq(w, h, 6, seed=qp, cut=cut)
    bs, recons, infos = _roundtrip(oracle, frames, w, h, qp=qp, merange=R, gop=32, cabac=1, deblock=deblock)
    types = collections.Counter()
    for inf in infos[1:]:
        types.update(inf["mb_type"].tolist())
    assert types[0] > 0
    if cut is not None:
        assert types[1] + types[2] > 0


def test_cabac_synthetic_pan_two_gops_and_size(oracle):
    """bench content: skip runs, I16x16 + I4x4 + inter; CABAC must decode identically and be smaller than CAVLC"""
    w, h = 192, 112
    frames = [oracle.synth_frame(w, h, t) for t in range(8)]
    bs_v, _, _ = _roundtrip(oracle, frames, w, h, qp=30, merange=16, gop=4, cabac=0)
    bs_a, _, _ = _roundtrip(oracle, frames, w, h, qp=30, merange=16, gop=4, cabac=1)
    assert len(bs_a) < len(bs_v)


def test_cabac_noise_large_levels(oracle):
    """saturated noise at low QP: long prefixes, Exp-Golomb escapes of coeff_abs_level_minus1 and of mvd"""
    rng = np.random.default_rng(7)
    w, h = 64, 48
    frames = [(rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8),
               rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)) for _ in range(3)]
    _roundtrip(oracle, frames, w, h, qp=10, merange=16, gop=32, cabac=1)


def test_sps_profile_bytes(oracle):
    """the reference reads profile_idc / compat / level straight from the SPS NAL (av_encode.c:703-705)"""
    for cabac, t8, prof in ((0, 0, 66), (1, 0, 77), (1, 1, 100)):
        sps = oracle.Entropy(320, 240, 26, cabac=cabac, transform8x8=t8).sps()
        assert sps[0] == 0x67 and sps[1] == prof and sps[3] >= 13


# ---- row N1: adaptive 8x8 transform of inter macroblocks (High profile), both entropy coders ------------------------
@pytest.mark.parametrize("cabac", [0, 1])
@pytest.mark.parametrize("w,h,qp,R,cut,deblock", [(176, 144, 26, 16, None, 1), (320, 240, 32, 32, 4, 1), (208, 160, 18, 16, 3, 0),
                                                  (318, 242, 28, 16, None, 1), (64, 48, 12, 16, 1, 0), (96, 80, 44, 16, 2, 1)])
def test_transform8x8_stream_decodes_to_oracle_recon(oracle, w, h, qp, R, cut, deblock, cabac):
    frames = smooth_seq(w, h, 6, seed=qp, cut=cut)
    bs, recons, infos = _roundtrip(oracle, frames, w, h, qp=qp, merange=R, gop=32, cabac=cabac, deblock=deblock, transform8x8=1)
    t8 = sum(int(inf["transform8x8"].sum()) for inf in infos)
    n4 = sum(int(((inf["mb_type"] == 0) & (inf["transform8x8"] == 0) & ((inf["cbp"] & 15) != 0)).sum()) for inf in infos)
    assert t8 > 0, "no macroblock chose the 8x8 transform"
    if qp < 40:
        assert n4 > 0, "no coded inter macroblock kept the 4x4 transform"


@pytest.mark.parametrize("cabac", [0, 1])
@pytest.mark.parametrize("w,h,qp,scale,deblock", [(96, 80, 40, 10, 1), (176, 144, 44, 12, 0), (208, 160, 36, 16, 1), (70, 54, 46, 8, 1)])
def test_intra8x8_stream_decodes_to_oracle_recon(oracle, w, h, qp, scale, deblock, cabac):
    """row N1, intra 8x8: reference-sample filter + nine 8x8 predictors + intra 8x8 residual, I and P slices"""
    frames = coarse_seq(w, h, 4, seed=qp, scale=scale)
    bs, recons, infos = _roundtrip(oracle, frames, w, h, qp=qp, merange=16, gop=2, cabac=cabac, deblock=deblock, transform8x8=1)
    i8 = [inf[inf["mb_type"] == 3] for inf in infos]
    assert sum(x.size for x in i8) >= 8, "too few I8x8 macroblocks to mean anything"
    for inf in infos:
        m = inf[inf["mb_type"] == 3]
        assert np.all(m["transform8x8"] == 1)
        assert np.all(m["i4_mode"].reshape(-1, 4, 4) == ((m["i8_modes"][:, None] >> (4 * np.arange(4))[None, :]) & 15)[:, :, None])


def test_intra8x8_all_nine_modes_are_exercised(oracle):
    w, h, qp = 96, 80, 40
    frames = coarse_seq(w, h, 4, seed=qp, scale=10)
    _, _, infos = _roundtrip(oracle, frames, w, h, qp=qp, merange=16, gop=2, cabac=1, deblock=1, transform8x8=1)
    modes = np.concatenate([((inf["i8_modes"][inf["mb_type"] == 3][:, None] >> (4 * np.arange(4))[None, :]) & 15).ravel() for inf in infos])
    assert set(modes.tolist()) == set(range(9))


# ---- row N1: inter partitions 16x8 / 8x16 / 8x8 (local refinement around the 16x16 vector) ---------------------------
from test_oracle_decode import shear_seq


@pytest.mark.parametrize("cabac,t8", [(0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("w,h,qp,deblock", [(176, 144, 26, 1), (320, 240, 30, 1), (208, 160, 20, 0), (96, 80, 36, 1), (70, 54, 30, 1)])
def test_partition_stream_decodes_to_oracle_recon(oracle, w, h, qp, deblock, cabac, t8):
    """mb_type / sub_mb_type, directional + median MV prediction per partition, mvd contexts, per-partition luma and
    chroma MC, boundary strengths from per-quadrant vectors: all pinned by the decoder round trip"""
    frames = shear_seq(w, h, 5, seed=qp)
    bs, recons, infos = _roundtrip(oracle, frames, w, h, qp=qp, merange=16, gop=32, cabac=cabac, deblock=deblock, transform8x8=t8,
                                   partitions=1)
    parts = np.bincount(np.concatenate([i["part"][i["mb_type"] == 0] for i in infos[1:]]), minlength=4)
    assert np.all(parts[1:] > 0), "every partition shape must occur: %s" % parts
    bs0, _, _ = _roundtrip(oracle, frames, w, h, qp=qp, merange=16, gop=32, cabac=cabac, deblock=deblock, transform8x8=t8)
    assert len(bs) < len(bs0)                                  # on sheared content partitions must pay off


def test_partitions_do_not_change_uniform_motion(oracle):
    """a pure pan has nothing to gain: the stream with partitions enabled stays the 16x16 stream"""
    w, h = 192, 112
    frames = [oracle.synth_frame(w, h, t) for t in range(4)]
    a, _, infos = _roundtrip(oracle, frames, w, h, qp=28, merange=16, gop=32, cabac=1, partitions=1)
    b, _, _ = _roundtrip(oracle, frames, w, h, qp=28, merange=16, gop=32, cabac=1, partitions=0)
    frac = np.mean(np.concatenate([i["part"][i["mb_type"] == 0] for i in infos[1:]]) != 0)
    assert frac < 0.2 and len(a) <= len(b) * 1.02


@pytest.mark.parametrize("cabac,t8", [(0, 0), (1, 1)])
@pytest.mark.parametrize("w,h,qp,amp", [(176, 144, 26, 5), (208, 160, 22, 3), (96, 80, 34, 7)])
def test_wide_partition_search_decodes_and_beats_local(oracle, w, h, qp, amp, cabac, t8):
    """partitions = 2: every part has its own exhaustive full-pel search.  With shears of several pixels per frame the local
    refinement (partitions = 1, +-3/4 pel around the 16x16 vector) cannot follow the parts; the wide search must"""
    frames = shear_seq(w, h, 4, seed=qp, amp=amp)
    sizes = {}
    for pm in (1, 2):
        bs, recons, infos = _roundtrip(oracle, frames, w, h, qp=qp, merange=16, gop=32, cabac=cabac, deblock=1, transform8x8=t8, partitions=pm)
        sizes[pm] = len(bs)
        if pm == 2:
            parts = np.bincount(np.concatenate([i["part"][i["mb_type"] == 0] for i in infos[1:]]), minlength=4)
            assert np.all(parts[1:] > 0), parts
            spread = max(int(np.abs(i["mv8"][:, 2, 0].astype(int) - i["mvx"]).max()) for i in infos[1:])
            assert spread > 8, "parts never moved more than 2 px apart: %d" % spread
    assert sizes[2] < sizes[1]


@pytest.mark.parametrize("cabac", [0, 1])
def test_all_features_extremes_decode(oracle, cabac):
    """one-macroblock pictures, saturated noise, flat and ramp content at QP 10 and 51 with every tool on"""
    rng = np.random.default_rng(23)
    for (w, h) in ((16, 16), (32, 16), (16, 48), (70, 38)):
        cw, ch = (w + 1) // 2, (h + 1) // 2
        noise = [((rng.integers(0, 2, (h, w)) * 255).astype(np.uint8), (rng.integers(0, 2, (ch, cw)) * 255).astype(np.uint8),
                  (rng.integers(0, 2, (ch, cw)) * 255).astype(np.uint8)) for _ in range(3)]
        ramp = [((np.add.outer(np.arange(h) * 3, np.arange(w) * 2) + 5 * t).astype(np.uint8), np.full((ch, cw), 100 + t, np.uint8),
                 np.full((ch, cw), 140 - t, np.uint8)) for t in range(3)]
        for frames in (noise, ramp):
            for qp in (10, 51):
                for pm in (1, 2):
                    _roundtrip(oracle, frames, w, h, qp=qp, merange=16, gop=32, cabac=cabac, deblock=1, transform8x8=1, partitions=pm)
