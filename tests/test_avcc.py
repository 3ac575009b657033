"""Row N3, wire format towards the muxer: b2_avcc_write's AVCDecoderConfigurationRecord plus
4-byte length-prefixed NAL samples (what av_encode.c:683-744 hands to libmp4v2) must form an MP4 that libavformat
demuxes and libavcodec decodes to the oracle's reconstruction.  CPU only (oracle decisions + the product's host writers)."""
import ctypes as C
import os
import struct
import numpy as np
import pytest
from test_oracle_decode import smooth_seq
from mp4mini import write_mp4


def _nals(annexb):
    return [p for p in annexb.split(b"\x00\x00\x00\x01")[1:]]


@pytest.mark.parametrize("cabac", [0, 1])
def test_avcc_and_length_prefixed_samples_make_a_decodable_mp4(oracle, b2, tmp_path, cabac):
    import cv2
    w, h, n = 176, 144, 6
    frames = smooth_seq(w, h, n, seed=3)
    bs, recons, infos, _ = oracle.encode_sequence(frames, w, h, qp=28, merange=16, gop=3, cabac=cabac)
    nals = _nals(bs)
    sps = next(x for x in nals if x[0] & 31 == 7); pps = next(x for x in nals if x[0] & 31 == 8)
    L = b2.lib()
    L.b2_avcc_write.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_int]
    buf = (C.c_uint8 * 256)()
    k = L.b2_avcc_write(sps, len(sps), pps, len(pps), buf, 256)
    assert k == 11 + len(sps) + len(pps)
    avcc = bytes(buf[:k])
    assert avcc[0] == 1 and avcc[1:4] == sps[1:4] and avcc[4] == 0xff and avcc[5] == 0xe1       # lengthSizeMinusOne = 3, 1 SPS
    assert L.b2_avcc_write(sps, len(sps), pps, len(pps), buf, 8) < 0                            # cap too small
    assert L.b2_avcc_write(pps, len(pps), sps, len(sps), buf, 256) < 0                          # swapped NAL types
    # one sample per picture: its slice NAL with a 4-byte big-endian length (SPS/PPS live in avcC, av_encode.c:722-727)
    samples = [struct.pack(">I", len(x)) + x for x in nals if x[0] & 31 in (1, 5)]
    sync = [x[0] & 31 == 5 for x in nals if x[0] & 31 in (1, 5)]
    assert len(samples) == n and sum(sync) == 2
    path = str(tmp_path / "t.mp4")
    write_mp4(path, avcc, samples, sync, w, h)
    cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
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
        y = fr.reshape(-1, w)[:h] if fr.ndim == 2 else fr[:h, :w, 0]
        assert np.array_equal(y, recons[i].y[:h, :w]), "frame %d differs after MP4 round trip" % i
