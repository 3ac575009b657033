"""Host logic of the drop-in encoder (video-encoder_b200/host/b2h_encoder.c, b2h_sws.c) on the CPU: the product's host
sources are linked against tests/mock/mock_engine.c -- an oracle-backed stand-in for the CUDA engine, test infrastructure
only -- so that GOP streaming, the per-GPU threads, the entropy worker pool, the display-order fifo, the x264 delay / flush
contract (av_encode.c:971-974, :1076-1083), the deferred sws_scale hand-over (:545-547 -> :970) and closed-GOP sharding over
several (pretend) GPUs run in the `-m "not gpu"` suite.  The same assertions run against the real library in
tests/test_dropin.py / tests/test_multi_gpu.py (-m gpu)."""
import ctypes as C
import glob
import os
import subprocess
import time
import numpy as np
import pytest
from test_oracle_decode import smooth_seq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mock(b2mod):
    out = os.path.join(ROOT, "tests", "mock", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libb2enc_mock.so")
    srcs = ([os.path.join(ROOT, "tests", "mock", "mock_engine.c")] + sorted(glob.glob(os.path.join(ROOT, "video-encoder_b200", "host", "*.c")))
            + sorted(glob.glob(os.path.join(ROOT, "oracle", "b2o_*.c"))))
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs + glob.glob(os.path.join(ROOT, "include", "*.h"))
                                     + glob.glob(os.path.join(ROOT, "video-encoder_b200", "host", "*.h"))):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-Wall", "-I" + os.path.join(ROOT, "include"),
                               "-I" + os.path.join(ROOT, "video-encoder_b200", "host"), "-o", so] + srcs + ["-lm", "-lpthread"])
    return C.CDLL(so)


@pytest.fixture(scope="module")
def b2mod():
    import b2enc
    return b2enc


def drive(b2, lib, frames, w, h, devices=1, check=None, **kw):
    os.environ["B2_MOCK_DEVICES"] = str(max(devices, 1))
    enc = b2.DropInEncoder(w, h, library=lib, i_devices=devices, **kw)
    out = []
    for t, fr in enumerate(frames):
        size, nals, pts, dts, key = enc.encode(fr, pts=1000 + 40 * t)
        assert size >= 0
        if size > 0:
            out.append((nals, pts, dts, key))
        assert enc.delayed() == t + 1 - len(out)                     # frames in - frames out, at every call
        if check:
            check(t, len(out))
    while enc.delayed() > 0:
        size, nals, pts, dts, key = enc.encode(None, 0)
        assert size > 0                                              # a flush call waits for its frame (av_encode.c:1076-1083)
        out.append((nals, pts, dts, key))
    assert enc.encode(None, 0)[0] == 0
    enc.close()
    return out


def annexb(out):
    return b"".join(d for nals, *_ in out for _, d in nals)


@pytest.mark.parametrize("gop,slots,n,cabac", [(4, 2, 11, 1), (3, 3, 13, 0), (5, 2, 5, 1), (4, 4, 1, 1)])
def test_stream_equals_oracle_encoder(oracle, b2mod, mock, gop, slots, n, cabac):
    """reference call sequence -> bitstream identical to the oracle encoder's, display order, pts passed through, key frames
    at GOP starts; tune film carries x264's deblock -1:-1 into the slice header and the loop filter"""
    w, h, qp = 64, 48, 30
    frames = smooth_seq(w, h, n, seed=5, cut=6)
    out = drive(b2mod, mock, frames, w, h, preset="medium", tune="film", quality=qp, profile=None if cabac else "baseline",
                annexb=1, i_keyint_max=gop, i_gop_slots=slots)
    assert len(out) == n
    assert [o[1] for o in out] == [1000 + 40 * t for t in range(n)]
    assert [o[3] for o in out] == [int(t % gop == 0) for t in range(n)]
    ref, *_ = oracle.encode_sequence(frames, w, h, qp=qp, merange=16, gop=gop, fps=(30, 1), deblock=1, cabac=cabac, deblock_offsets=(-1, -1))
    assert annexb(out) == ref


def test_first_output_does_not_wait_for_a_batch(b2mod, mock):
    """GOPs are encoded while they are gathered: the first frame comes back long before slots x keyint pictures are in"""
    w, h, gop, slots = 48, 32, 4, 4
    frames = smooth_seq(w, h, 40, seed=1)
    first = []

    def check(t, nout):
        if nout and not first:
            first.append(t)
        time.sleep(0.01)                                             # a producer slower than the (mock) GPU, e.g. a decoder
    out = drive(b2mod, mock, frames, w, h, quality=32, annexb=1, i_keyint_max=gop, i_gop_slots=slots, check=check)
    assert len(out) == 40
    assert first and first[0] < 2 * gop, "first output after %s pictures" % first


@pytest.mark.parametrize("n,gop,slots", [(23, 4, 2), (9, 2, 3)])
def test_gpu_count_does_not_change_the_stream(b2mod, mock, n, gop, slots):
    """T5 on the host logic: closed GOP k goes to GPU k % N; 1, 2 and 3 GPUs give byte-identical streams in display order"""
    w, h = 48, 48
    frames = smooth_seq(w, h, n, seed=9, cut=7)
    streams = []
    for devices in (1, 2, 3):
        out = drive(b2mod, mock, frames, w, h, devices=devices, quality=31, annexb=0, i_keyint_max=gop, i_gop_slots=slots)
        assert len(out) == n and [o[1] for o in out] == [1000 + 40 * t for t in range(n)]
        streams.append(annexb(out))
    assert streams[0] == streams[1] == streams[2]


def test_devices_from_environment(b2mod, mock):
    """an unmodified av_encode.c cannot set i_devices: B2ENC_DEVICES does; more GPUs than visible is an error (NULL handle)"""
    w, h = 32, 32
    frames = smooth_seq(w, h, 6, seed=3)
    one = annexb(drive(b2mod, mock, frames, w, h, quality=30, annexb=1, i_keyint_max=2, i_gop_slots=2))
    os.environ["B2ENC_DEVICES"] = "2"
    try:
        os.environ["B2_MOCK_DEVICES"] = "2"
        enc = b2mod.DropInEncoder(w, h, library=mock, quality=30, annexb=1, i_keyint_max=2, i_gop_slots=2)
        out = []
        for t, fr in enumerate(frames):
            r = enc.encode(fr, t)
            if r[0] > 0: out.append((r[1],))
        while enc.delayed() > 0:
            out.append((enc.encode(None, 0)[1],))
        enc.close()
        assert annexb(out) == one
        os.environ["B2_MOCK_DEVICES"] = "1"
        with pytest.raises(RuntimeError):
            b2mod.DropInEncoder(w, h, library=mock, quality=30, i_keyint_max=2, i_gop_slots=2)
    finally:
        del os.environ["B2ENC_DEVICES"]


def test_zero_latency(b2mod, mock):
    w, h = 48, 32
    frames = smooth_seq(w, h, 5, seed=2)
    os.environ["B2_MOCK_DEVICES"] = "1"
    enc = b2mod.DropInEncoder(w, h, library=mock, tune="zerolatency", quality=30, i_keyint_max=3)
    for t, fr in enumerate(frames):
        size, nals, pts, dts, key = enc.encode(fr, t)
        assert size > 0 and pts == t and key == int(t % 3 == 0) and enc.delayed() == 0
    enc.close()


def to_fmt(fmt, y, u, v):
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
def test_sws_scale_into_the_encoder_picture(oracle, b2mod, mock, fmts):
    """sws_scale(decoder picture -> pic_in) followed by encode(pic_in), the reference's per-frame pair (av_encode.c:545-547,
    :970): the conversion is deferred into the encoder; the stream equals the one of separately converted pictures.  A change
    of the decoder format mid-stream (third case) closes the GOP in flight and re-shapes the input rings."""
    w, h, gop, qp = 64, 32, 3, 30
    n = 6 * len(fmts)
    base = smooth_seq(w, h, n, seed=11)
    os.environ["B2_MOCK_DEVICES"] = "1"
    enc = b2mod.DropInEncoder(w, h, library=mock, quality=qp, annexb=1, i_keyint_max=gop, i_gop_slots=2)
    out, conv = [], []
    for t, (y, u, v) in enumerate(base):
        fmt = fmts[t // 6]
        planes = to_fmt(fmt, y, u, v)
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


def test_sws_host_output_flag_and_plain_destination(oracle, b2mod, mock):
    """B2_SWS_HOST_OUTPUT, or a destination that is not an encoder picture, gives sws_scale's host-out behaviour"""
    w, h = 32, 16
    y, u, v = smooth_seq(w, h, 1, seed=4)[0]
    planes = to_fmt("yuyv422", y, u, v)
    want = oracle.convert_to_i420("yuyv422", w, h, planes)
    os.environ["B2_MOCK_DEVICES"] = "1"
    enc = b2mod.DropInEncoder(w, h, library=mock, quality=30, i_keyint_max=2, i_gop_slots=2)
    enc.encode_via_sws("yuyv422", planes, 0, flags=1 | 0x10000000)
    got = [np.frombuffer((C.c_uint8 * (a.size)).from_address(enc.pic_in.img.plane[i]), np.uint8).reshape(a.shape) for i, a in enumerate(want)]
    assert all(np.array_equal(g, a) for g, a in zip(got, want))
    while enc.delayed() > 0:
        enc.encode(None, 0)
    enc.close()


def test_many_gops_backpressure(b2mod, mock):
    """more GOPs than slots x GPUs: the caller blocks for a free slot, nothing is lost or reordered"""
    w, h, n = 32, 32, 90
    frames = smooth_seq(w, h, 10, seed=6)
    frames = [frames[t % 10] for t in range(n)]
    out = drive(b2mod, mock, frames, w, h, devices=2, quality=34, annexb=1, i_keyint_max=3, i_gop_slots=2)
    assert len(out) == n and [o[1] for o in out] == [1000 + 40 * t for t in range(n)]
    single = drive(b2mod, mock, frames, w, h, devices=1, quality=34, annexb=1, i_keyint_max=3, i_gop_slots=5)
    assert annexb(out) == annexb(single)


def test_randomised_call_sequences(oracle, b2mod, mock):
    """random picture sizes, GOP lengths, slot counts, GPU counts, frame counts (partial last GOPs, fewer frames than one GOP,
    more GOPs than slots), entropy coder, container format and producer pacing: always the oracle encoder's bytes, in order"""
    rng = np.random.default_rng(20261018)
    for case in range(24):
        w, h = int(rng.integers(1, 5)) * 16, int(rng.integers(1, 4)) * 16
        gop = int(rng.integers(1, 7)); slots = int(rng.integers(2, 6)); devices = int(rng.integers(1, 4)); n = int(rng.integers(1, 40))
        cabac = int(rng.random() < 0.6); annexb_ = int(rng.random() < 0.5); qp = int(rng.integers(20, 45))
        frames = smooth_seq(w, h, n, seed=int(rng.integers(0, 1 << 30)), cut=(int(rng.integers(1, n)) if n > 2 and rng.random() < 0.3 else None))
        pace = float(rng.choice([0.0, 0.0, 0.002]))
        out = drive(b2mod, mock, frames, w, h, devices=devices, preset="medium", tune="film", quality=qp, profile=None if cabac else "baseline",
                    annexb=annexb_, i_keyint_max=gop, i_gop_slots=slots, check=(lambda t, k: time.sleep(pace)) if pace else None)
        assert len(out) == n and [o[1] for o in out] == [1000 + 40 * t for t in range(n)], case
        bs = b""
        for nals, *_ in out:
            for _, d in nals:
                bs += d if annexb_ else b"\x00\x00\x00\x01" + d[4:]
        ref, *_ = oracle.encode_sequence(frames, w, h, qp=qp, merange=16, gop=gop, fps=(30, 1), deblock=1, cabac=cabac, deblock_offsets=(-1, -1))
        assert bs == ref, "case %d: %dx%d gop %d slots %d devices %d frames %d" % (case, w, h, gop, slots, devices, n)


@pytest.mark.parametrize("fmt", ["yuyv422", "yuv420p", "nv12"])
def test_large_picture_staging_copy_is_threaded_and_exact(oracle, b2mod, mock, fmt):
    """pictures of 2 MB and more are staged by the caller plus helper threads (chunks of rows claimed one at a time, one
    post per picture, host/b2h_sws.c) into the double-buffered page-locked staging: strided 1080p sources (one, two and three
    planes), stream == oracle encoder's on the converted pictures"""
    w, h, n, gop = 1920, 1080, 3, 2
    base = [oracle.synth_frame(w, h, t, 3) for t in range(n)]
    os.environ["B2_MOCK_DEVICES"] = "1"
    os.environ["B2ENC_SWS_THREADS"] = "4"
    try:
        enc = b2mod.DropInEncoder(w, h, library=mock, preset="ultrafast", tune="film", quality=38, annexb=1, i_keyint_max=gop, i_gop_slots=2)
        out, conv = [], []
        for t, (y, u, v) in enumerate(base):
            planes = to_fmt(fmt, y, u, v)                         # padded row pitch: every row is a separate copy
            conv.append(oracle.convert_to_i420(fmt, w, h, planes))
            r = enc.encode_via_sws(fmt, planes, t)
            if r[0] > 0: out.append(r[1])
        while enc.delayed() > 0:
            out.append(enc.encode(None, 0)[1])
        enc.close()
    finally:
        del os.environ["B2ENC_SWS_THREADS"]
    assert len(out) == n
    # preset ultrafast: no sub-pel, no intra in P, CAVLC, loop filter off
    ref, *_ = oracle.encode_sequence(conv, w, h, qp=38, merange=16, subpel=0, intra_in_p=0, gop=gop, fps=(30, 1), deblock=0, cabac=0)
    assert b"".join(d for nals in out for _, d in nals) == ref


def test_staging_copy_pool_stress():
    """the helper pool of the staging copy (host/b2h_sws.c: one post per picture, helpers poll before they sleep) on random plane shapes
    and pitches with pauses that let the helpers fall asleep: every picture's contents exact (scripts/sws_copy_stress.c)"""
    import subprocess, tempfile
    exe = os.path.join(tempfile.mkdtemp(prefix="b2_swscopy_"), "stress")
    subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "video-encoder_b200", "host"),
                           "-o", exe, os.path.join(ROOT, "scripts", "sws_copy_stress.c"), "-lpthread"])
    for helpers in ("3", "1", "0"):
        r = subprocess.run([exe, "1500", helpers], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "contents exact" in r.stdout, r.stdout + r.stderr


def test_stats_and_host_ceiling_harness(tmp_path):
    """B2ENC_STATS=1 reports where the host threads spent their time, and scripts/host_ceiling.py (zero-latency mock engine,
    B2_MOCK_CANNED) carries a small stream end to end: every frame in comes out"""
    import subprocess, sys
    env = dict(os.environ, B2ENC_STATS="1", B2_CEILING_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "host_ceiling.py"), "320", "192", "200", "2", "3"], capture_output=True, text=True,
                       env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "200 frames in, 200 out" in r.stdout and "2 pretend GPU(s), 3 GOP slots each" in r.stdout
    assert r.stderr.count("b2enc stats: GPU thread") == 2 and "entropy workers" in r.stderr
