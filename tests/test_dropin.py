"""Drop-in boundary on the GPU: the b2_* mirror of the x264/swscale calls made by av_encode.c, driven with
the reference's own call sequence; ABI contracts of SURVEY.md 8b (T6), bitstream identical to the oracle
encoder's, decodable, N-slot output == 1-slot output (T5)."""
import os
import numpy as np
import pytest
from test_oracle_decode import smooth_seq, shear_seq

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sws_golden.npz"))
CASES = sorted({k.rsplit("_", 1)[0] for k in G.files})


@pytest.mark.parametrize("case", CASES)
def test_sws_scale_matches_live_swscale_golden(b2, case):
    fmt, size = case.split("_")
    w, h = [int(a) for a in size.split("x")]
    ins = [G[f"{case}_in{i}"] for i in range(3) if f"{case}_in{i}" in G.files]
    y, u, v = b2.sws_convert(fmt, w, h, ins, dst_pad=5)
    assert np.array_equal(y, G[f"{case}_out0"]) and np.array_equal(u, G[f"{case}_out1"]) and np.array_equal(v, G[f"{case}_out2"])


T = np.load(os.path.join(os.path.dirname(__file__), "golden", "sws_tolerance.npz"))
TCASES = sorted({k.rsplit("_", 1)[0] for k in T.files})


@pytest.mark.parametrize("case", TCASES)
def test_sws_scale_tolerance_formats(oracle, b2, case):
    """row N4: rgb24 / yuv422p / yuv411p through b2_sws_scale (K0): identical to the oracle's drift-free closed form, and
    within the pinned tolerance of the live libswscale output"""
    fmt, size = case.split("_")
    w, h = [int(a) for a in size.split("x")]
    ins = [T[f"{case}_in{i}"] for i in range(3) if f"{case}_in{i}" in T.files]
    y, u, v = b2.sws_convert(fmt, w, h, ins, dst_pad=3)
    oy, ou, ov = oracle.convert_to_i420(fmt, w, h, ins)
    assert np.array_equal(y, oy) and np.array_equal(u, ou) and np.array_equal(v, ov)
    tol = {"rgb24": (1, 1), "yuv422p": (1, 1), "yuv411p": (1, 2)}[fmt]
    assert np.abs(y.astype(int) - T[f"{case}_out0"]).max() <= tol[0]
    assert np.abs(u.astype(int) - T[f"{case}_out1"]).max() <= tol[1] and np.abs(v.astype(int) - T[f"{case}_out2"]).max() <= tol[1]


def test_sws_refuses_sizes_without_closed_form(b2):
    L = b2._dropin_lib()
    assert not L.b2_sws_getContext(70, 38, b2.FMT["yuv411p"], 70, 38, 0, 1, None, None, None)     # 4:1:1 needs w % 4 == 0
    assert not L.b2_sws_getContext(61, 36, b2.FMT["bgr24"], 61, 36, 0, 1, None, None, None)       # odd width
    assert not L.b2_sws_getContext(64, 35, b2.FMT["rgb24"], 64, 35, 0, 1, None, None, None)       # odd height
    assert not L.b2_sws_getContext(64, 36, 99, 64, 36, 0, 1, None, None, None)


def drive(b2, frames, w, h, **kw):
    """the reference loop: encode every frame (av_encode.c:968-975), then drain (:1076-1083)"""
    enc = b2.DropInEncoder(w, h, **kw)
    out = []
    for t, fr in enumerate(frames):
        size, nals, pts, dts, key = enc.encode(fr, pts=1000 + 40 * t)
        assert size >= 0
        if size > 0:
            out.append((nals, pts, dts, key))
    guard = 0
    while enc.delayed() > 0:
        size, nals, pts, dts, key = enc.encode(None, 0)
        assert size > 0
        out.append((nals, pts, dts, key))
        guard += 1
        assert guard <= len(frames)
    assert enc.encode(None, 0)[0] == 0
    enc.close()
    return out


def to_annexb(out, length_prefixed):
    bs = b""
    for nals, *_ in out:
        for _, data in nals:
            if length_prefixed:
                assert int.from_bytes(data[:4], "big") == len(data) - 4          # av_encode.c:722
                bs += b"\x00\x00\x00\x01" + data[4:]
            else:
                assert data[:4] == b"\x00\x00\x00\x01"
                bs += data
    return bs


@pytest.mark.parametrize("profile,profile_idc,cabac,t8", [(None, 77, 1, 0), ("baseline", 66, 0, 0), ("high", 100, 1, 1)])
def test_reference_call_sequence_and_bitstream(oracle, b2, profile, profile_idc, cabac, t8):
    """profile NULL is what the reference passes by default (av_encode.c:105, :403): x264's preset default = CABAC"""
    w, h, qp, gop, n = 176, 144, 27, 4, 11
    frames = smooth_seq(w, h, n, seed=4, cut=6)
    out = drive(b2, frames, w, h, preset="medium", tune="film", quality=qp, profile=profile, i_keyint_max=gop, i_gop_slots=2,
                b_transform_8x8=t8)
    assert len(out) == n
    assert [o[1] for o in out] == [1000 + 40 * t for t in range(n)]               # display order, pts passed through
    assert [o[3] for o in out] == [int(t % gop == 0) for t in range(n)]           # keyframe flag (av_encode.c:783)
    sps = out[0][0][0]
    assert sps[0] == 7 and out[0][0][1][0] == 8 and out[0][0][2][0] == 5          # SPS, PPS, IDR slice (av_encode.c:683-736)
    assert sps[1][5] == profile_idc and sps[1][7] in (10, 11, 12, 13, 21, 22, 30)                  # profile_idc / level_idc at [5],[7] (:703-705)
    bs = to_annexb(out, length_prefixed=True)
    ref_bs, recons, _, _ = oracle.encode_sequence(frames, w, h, qp=qp, merange=16, gop=gop, fps=(30, 1), deblock=1, cabac=cabac,
                                                   transform8x8=t8, deblock_offsets=(-1, -1))   # tune film: deblock -1:-1 as in x264
    assert bs == ref_bs, "GPU drop-in bitstream differs from the oracle encoder's"
    dec = oracle.decode_yuv(oracle.split_access_units(bs))
    assert len(dec) == n
    for i, (dy, du, dv) in enumerate(dec):
        assert np.array_equal(dy, recons[i].y[:h, :w]) and np.array_equal(du, recons[i].u[:h // 2, :w // 2])


@pytest.mark.parametrize("pm,amp", [(1, 1), (2, 4)])
def test_dropin_high_profile_with_partitions(oracle, b2, pm, amp):
    """everything of row N1 at once through the x264-mirror: CABAC + 8x8 transform + intra 8x8 + inter partitions
    (1: local refinement, 2: own full-pel search per part)"""
    w, h, qp, gop, n = 176, 144, 28, 5, 10
    frames = shear_seq(w, h, n, seed=12, amp=amp)
    out = drive(b2, frames, w, h, preset="medium", tune="film", quality=qp, profile="high", i_keyint_max=gop, i_gop_slots=2,
                b_transform_8x8=1, b_partitions=pm)
    bs = to_annexb(out, length_prefixed=True)
    ref_bs, recons, infos, _ = oracle.encode_sequence(frames, w, h, qp=qp, merange=16, gop=gop, fps=(30, 1), deblock=1, cabac=1,
                                                      transform8x8=1, partitions=pm, deblock_offsets=(-1, -1))
    assert bs == ref_bs
    assert sum(int((i["part"] != 0).sum()) for i in infos) > 20
    dec = oracle.decode_yuv(oracle.split_access_units(bs))
    assert len(dec) == n and all(np.array_equal(d[0], r.y[:h, :w]) for d, r in zip(dec, recons))


@pytest.mark.parametrize("fmt", ["nv12", "yuyv422", "bgr24"])
def test_decoder_format_straight_into_the_encoder(oracle, b2, fmt):
    """b2_param_t.i_csp_in: pictures in the decoder's own format (strided, pageable planes) go to b2_encoder_encode without an
    sws_scale call; the conversion runs as the engine's first kernel and the stream equals the one of converted I420 pictures"""
    w, h, n, gop, qp = 96, 64, 7, 3, 29
    rng = np.random.default_rng(3)
    base = smooth_seq(w, h, n, seed=21)
    raws, conv = [], []
    for (y, u, v) in base:
        if fmt == "nv12":
            uv = np.empty((h // 2, w), np.uint8); uv[:, 0::2] = u; uv[:, 1::2] = v
            pl = [np.pad(y, ((0, 0), (0, 8))), np.pad(uv, ((0, 0), (0, 16)))]
        elif fmt == "yuyv422":
            p = np.empty((h, 2 * w), np.uint8); p[:, 0::2] = y
            p[:, 1::4] = np.repeat(u, 2, axis=0); p[:, 3::4] = np.repeat(v, 2, axis=0)
            pl = [np.pad(p, ((0, 0), (0, 12)))]
        else:
            p = rng.integers(0, 256, (h, 3 * w), dtype=np.uint8)
            p[:, 0::3] = y; p[:, 1::3] = y // 2 + 40                     # correlated channels: something to predict
            pl = [np.pad(p, ((0, 0), (0, 9)))]
        raws.append(pl)
        conv.append(oracle.convert_to_i420(fmt, w, h, pl))
    enc = b2.DropInEncoder(w, h, preset="medium", tune="film", quality=qp, annexb=1, i_keyint_max=gop, i_gop_slots=2,
                           i_csp_in=b2.FMT[fmt])
    out = []
    for t, pl in enumerate(raws):
        size, nals, *_ = enc.encode_raw(pl, t)
        if size > 0: out.append((nals,))
    while enc.delayed() > 0:
        size, nals, *_ = enc.encode(None, 0)
        out.append((nals,))
    enc.close()
    assert len(out) == n
    bs = to_annexb(out, length_prefixed=False)
    ref_bs, recons, _, _ = oracle.encode_sequence(conv, w, h, qp=qp, merange=16, gop=gop, fps=(30, 1), deblock=1, cabac=1,
                                                   deblock_offsets=(-1, -1))
    assert bs == ref_bs


def test_slot_count_does_not_change_the_stream(b2):
    """closed GOPs are independent: 1, 3 and 4 GOPs in flight give byte-identical output (T5)"""
    w, h, n = 128, 96, 13
    frames = smooth_seq(w, h, n, seed=8)
    streams = []
    for slots in (1, 3, 4):
        out = drive(b2, frames, w, h, preset="slow", tune=None, quality=30, annexb=1, i_keyint_max=3, i_gop_slots=slots)
        assert len(out) == n
        streams.append(to_annexb(out, length_prefixed=False))
    assert streams[0] == streams[1] == streams[2]


def test_search_pruning_does_not_change_the_stream(b2):
    """b_me_prune (default 1: successive elimination in front of the exhaustive search) is lossless: the stream is byte-identical
    with the switch off, +-16 and +-32, with a scene cut so that predictors are poor"""
    w, h, n = 176, 144, 9
    frames = smooth_seq(w, h, n, seed=21, cut=4)
    for preset in ("medium", "slow"):
        streams = []
        for prune in (0, 1):
            out = drive(b2, frames, w, h, preset=preset, tune=None, quality=28, annexb=1, i_keyint_max=5, i_gop_slots=2, b_me_prune=prune)
            assert len(out) == n
            streams.append(to_annexb(out, length_prefixed=False))
        assert streams[0] == streams[1], preset


def test_delay_contract(b2):
    """0 = no output yet, delayed_frames() = frames in - frames out at every call, encode(NULL) returns one frame per call until
    the encoder is empty (av_encode.c:971-974, :1076-1083); GOPs are encoded while they are gathered, so the first frame
    arrives long before slots x keyint pictures are in"""
    import time
    w, h, gop = 64, 64, 4
    frames = smooth_seq(w, h, 24, seed=2)
    enc = b2.DropInEncoder(w, h, quality=30, i_keyint_max=gop, i_gop_slots=3)
    got, first = 0, None
    for t, fr in enumerate(frames):
        size = enc.encode(fr, t)[0]
        assert size >= 0
        got += size > 0
        if size > 0 and first is None: first = t
        assert enc.delayed() == t + 1 - got
        time.sleep(0.005)                                # a producer slower than the GPU (a decoder)
    assert first is not None and first < 2 * gop, "first output after %s pictures" % first
    while enc.delayed() > 0:
        assert enc.encode(None, 0)[0] > 0
        got += 1
    assert got == len(frames) and enc.encode(None, 0)[0] == 0
    enc.close()
    enc = b2.DropInEncoder(w, h, tune="zerolatency", quality=30)
    for t, fr in enumerate(frames[:5]):
        assert enc.encode(fr, t)[0] > 0 and enc.delayed() == 0
    enc.close()


def _to_fmt(fmt, y, u, v):
    h, w = y.shape
    if fmt == "yuv420p":
        return [np.pad(y, ((0, 0), (0, 8))), np.pad(u, ((0, 0), (0, 4))), np.pad(v, ((0, 0), (0, 4)))]
    if fmt == "nv12":
        uv = np.empty((h // 2, w), np.uint8); uv[:, 0::2] = u; uv[:, 1::2] = v
        return [np.pad(y, ((0, 0), (0, 8))), np.pad(uv, ((0, 0), (0, 16)))]
    p = np.empty((h, 2 * w), np.uint8); p[:, 0::2] = y
    p[:, 1::4] = np.repeat(u, 2, axis=0); p[:, 3::4] = np.repeat(v, 2, axis=0)
    return [np.pad(p, ((0, 0), (0, 12)))]


@pytest.mark.parametrize("fmts", [["yuv420p"], ["yuyv422"], ["nv12", "yuv420p", "yuyv422"]])
def test_sws_scale_into_the_encoder_picture(oracle, b2, fmts):
    """the reference's per-frame pair sws_scale(decoder picture -> pic_in), x264_encoder_encode(pic_in) (av_encode.c:545-547,
    :970): the conversion is deferred into the encoder (one PCIe crossing, K0 on the encoder's own planes) and the stream equals
    the one of separately converted pictures; a decoder-format change mid-stream re-shapes the input rings"""
    w, h, gop, qp = 176, 144, 3, 30
    n = 6 * len(fmts)
    base = smooth_seq(w, h, n, seed=11)
    enc = b2.DropInEncoder(w, h, quality=qp, annexb=1, i_keyint_max=gop, i_gop_slots=2)
    out, conv = [], []
    for t, (y, u, v) in enumerate(base):
        fmt = fmts[t // 6]
        planes = _to_fmt(fmt, y, u, v)
        conv.append(oracle.convert_to_i420(fmt, w, h, planes))
        r = enc.encode_via_sws(fmt, planes, t)
        if r[0] > 0: out.append((r[1], r[2]))
    while enc.delayed() > 0:
        r = enc.encode(None, 0)
        out.append((r[1], r[2]))
    enc.close()
    assert [o[1] for o in out] == list(range(n))
    ref, *_ = oracle.encode_sequence(conv, w, h, qp=qp, merange=16, gop=gop, fps=(30, 1), deblock=1, cabac=1, deblock_offsets=(-1, -1))
    assert b"".join(d for nals, _ in out for _, d in nals) == ref


def test_sws_host_output_flag(oracle, b2):
    """B2_SWS_HOST_OUTPUT: the encoder picture's planes are written (GPU round trip through K0) even though it is a registered picture"""
    import ctypes as C
    w, h = 64, 48
    y, u, v = smooth_seq(w, h, 1, seed=4)[0]
    planes = _to_fmt("yuyv422", y, u, v)
    want = oracle.convert_to_i420("yuyv422", w, h, planes)
    enc = b2.DropInEncoder(w, h, quality=30, i_keyint_max=2, i_gop_slots=2)
    enc.encode_via_sws("yuyv422", planes, 0, flags=1 | 0x10000000)
    got = [np.frombuffer((C.c_uint8 * (a.size)).from_address(enc.pic_in.img.plane[i]), np.uint8).reshape(a.shape).copy() for i, a in enumerate(want)]
    while enc.delayed() > 0:
        enc.encode(None, 0)
    enc.close()
    assert all(np.array_equal(g, a) for g, a in zip(got, want))
