"""a1 conversion: the oracle's closed forms against golden vectors produced by the LIVE libswscale
9.1.100 (tests/golden/make_sws_golden.py; same flags as av_encode.c:427-430).  CPU only."""
import os
import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sws_golden.npz"))
CASES = sorted({k.rsplit("_", 1)[0] for k in G.files})


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_swscale(oracle, case):
    fmt, size = case.split("_")
    w, h = [int(a) for a in size.split("x")]
    ins = [G[f"{case}_in{i}"] for i in range(3) if f"{case}_in{i}" in G.files]
    y, u, v = oracle.convert_to_i420(fmt, w, h, ins)
    assert np.array_equal(y, G[f"{case}_out0"])
    assert np.array_equal(u, G[f"{case}_out1"])
    assert np.array_equal(v, G[f"{case}_out2"])


# ---- row N4: formats that libswscale runs through its drifting fast-bilinear scaler -- tolerance pin ----
T = np.load(os.path.join(os.path.dirname(__file__), "golden", "sws_tolerance.npz"))
TCASES = sorted({k.rsplit("_", 1)[0] for k in T.files})
TOL = {"rgb24": (1, 1), "yuv422p": (1, 1), "yuv411p": (1, 2)}          # (luma, chroma) max abs difference


@pytest.mark.parametrize("case", TCASES)
def test_oracle_close_to_swscale_on_tolerance_formats(oracle, case):
    fmt, size = case.split("_")
    w, h = [int(a) for a in size.split("x")]
    ins = [T[f"{case}_in{i}"] for i in range(3) if f"{case}_in{i}" in T.files]
    y, u, v = oracle.convert_to_i420(fmt, w, h, ins)
    ty, tc = TOL[fmt]
    assert np.abs(y.astype(int) - T[f"{case}_out0"]).max() <= ty
    assert np.abs(u.astype(int) - T[f"{case}_out1"]).max() <= tc
    assert np.abs(v.astype(int) - T[f"{case}_out2"]).max() <= tc
    # and on average the conversion is unbiased (no systematic offset against the library)
    assert abs(float((y.astype(int) - T[f"{case}_out0"]).mean())) < 0.2
