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
    ref_bs, recons, _, _ = oracle.encode_sequence(frames, w, h, qp=24, merange=16, gop=4, fps=(30, 1), deblock=1, cabac=1)
    assert bs == ref_bs                                  # same stream as the oracle encoder at the same settings
    dec = oracle.decode_yuv(oracle.split_access_units(bs))
    assert len(dec) == n
    src = oracle.OFrame(w, h).load(*frames[-1])
    psnr = oracle.lib().b2o_psnr_y(oracle.C.byref(src.f), oracle.C.byref(recons[-1].f))
    assert psnr > 35.0
    # error convention: bad preset -> non-zero exit, message on stderr (av_encode.c:384-386)
    r = subprocess.run([exe, "--preset", "warp9", str(y4m), str(out)], capture_output=True, text=True)
    assert r.returncode != 0 and "preset" in r.stderr
