"""tools/b2_encode (host driver in C over include/b2enc.h, reference call sequence and command line) on the GPU."""
import os
import subprocess
import numpy as np
import pytest
from test_oracle_decode import smooth_seq

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_y4m_to_h264(oracle, tmp_path):
    exe = os.path.join(ROOT, "tools", "b2_encode")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "video-encoder_b200"), "-s"])
    w, h, n = 176, 144, 9
    frames = smooth_seq(w, h, n, seed=11)
    y4m = tmp_path / "in.y4m"; out = tmp_path / "out.h264"
    with open(y4m, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420jpeg\n" % (w, h))
        for y, u, v in frames:
            f.write(b"FRAME\n" + y.tobytes() + u.tobytes() + v.tobytes())
    r = subprocess.run([exe, "--preset", "medium", "--tune", "film", "--quality", "24", "--gop", "4", "--slots", "2",
                        str(y4m), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "9 frames in, 9 frames out" in r.stdout
    bs = open(out, "rb").read()
    ref_bs, recons, _, _ = oracle.encode_sequence(frames, w, h, qp=24, merange=16, gop=4, fps=(30, 1), deblock=1, cabac=1,
                                                   deblock_offsets=(-1, -1))      # tune film = deblock -1:-1, as in x264
    assert bs == ref_bs                                  # same stream as the oracle encoder at the same settings
    dec = oracle.decode_yuv(oracle.split_access_units(bs))
    assert len(dec) == n
    src = oracle.OFrame(w, h).load(*frames[-1])
    psnr = oracle.lib().b2o_psnr_y(oracle.C.byref(src.f), oracle.C.byref(recons[-1].f))
    assert psnr > 35.0
    # error convention: bad preset -> non-zero exit, message on stderr (av_encode.c:384-386)
    r = subprocess.run([exe, "--preset", "warp9", str(y4m), str(out)], capture_output=True, text=True)
    assert r.returncode != 0 and "preset" in r.stderr


def test_cli_filters_and_mp4(oracle, tmp_path):
    """the reference's command line with --filters (av_encode.c:116) and an .mp4 output: GPU pre-filters -> conversion -> encoder ->
    length-prefixed NALs + avcC in an MP4 that libavformat demuxes; decoded luma == oracle pipeline (filters, encoder) recon"""
    import cv2
    exe = os.path.join(ROOT, "tools", "b2_encode")
    w, h, n = 176, 144, 9
    frames = smooth_seq(w, h, n, seed=13)
    y4m = tmp_path / "in.y4m"; out = tmp_path / "out.mp4"
    with open(y4m, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F30:1 It A1:1 C420jpeg\n" % (w, h))
        for y, u, v in frames:
            f.write(b"FRAME\n" + y.tobytes() + u.tobytes() + v.tobytes())
    r = subprocess.run([exe, "--preset", "medium", "--tune", "film", "--quality", "24", "--gop", "4", "--slots", "2", "--profile", "high",
                        "--8x8dct", "--partitions", "2", "--filters", "hqdn3d,yadif", str(y4m), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "9 frames in, 9 frames out" in r.stdout
    dn = oracle.Hqdn3d(w, h)
    filt = oracle.yadif_sequence([dn(f) for f in frames], w, h, tff=1)
    _, recons, _, _ = oracle.encode_sequence(filt, w, h, qp=24, merange=16, gop=4, fps=(30, 1), deblock=1, cabac=1, transform8x8=1, partitions=2,
                                            deblock_offsets=(-1, -1))
    cap = cv2.VideoCapture(str(out), cv2.CAP_FFMPEG)
    cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    got = []
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        got.append(fr.copy())
    cap.release()
    assert len(got) == n
    for i, fr in enumerate(got):
        yy = fr.reshape(-1, w)[:h] if fr.ndim == 2 else fr[:h, :w, 0]
        assert np.array_equal(yy, recons[i].y[:h, :w]), "frame %d of the MP4 differs from the oracle pipeline" % i
    r = subprocess.run([exe, "--filters", "unsharp", str(y4m), str(out)], capture_output=True, text=True)
    assert r.returncode != 0 and "filter" in r.stderr
