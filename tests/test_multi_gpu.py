"""Row X2 / SURVEY.md 8e, T5: ONE stream spread over the GPUs of the box by closed GOP, inside libb2enc.so
(b2_param_t.i_devices: GOP k -> device i_device + k % N, one host thread + engine per GPU, frames returned in display order).
The N-GPU stream must be byte-identical to the 1-GPU stream.  Needs >= 2 GPUs (`gpurun --gpus N`); the same host logic runs on
the CPU against the mock engine in tests/test_host_pipeline.py."""
import hashlib
import numpy as np
import pytest
from test_oracle_decode import smooth_seq
from test_dropin import drive, to_annexb

pytestmark = pytest.mark.gpu


def device_counts(b2):
    n = b2.lib().b2_device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    return [k for k in (1, 2, 4, 8) if k <= n]


def test_gpu_count_does_not_change_the_stream_vs_oracle(oracle, b2):
    """small pictures, every tool on: 1..N GPUs give the oracle encoder's bytes"""
    w, h, qp, gop, n = 176, 144, 28, 3, 26
    frames = smooth_seq(w, h, n, seed=21, cut=11)
    ref, *_ = oracle.encode_sequence(frames, w, h, qp=qp, merange=32, gop=gop, fps=(30, 1), deblock=1, cabac=1, transform8x8=1,
                                     partitions=2, deblock_offsets=(-1, -1))
    for devices in device_counts(b2):
        out = drive(b2, frames, w, h, preset="slow", tune="film", quality=qp, profile="high", annexb=1, i_keyint_max=gop,
                    i_gop_slots=3, b_transform_8x8=1, b_partitions=2, i_devices=devices)
        assert [o[1] for o in out] == [1000 + 40 * t for t in range(n)]
        assert to_annexb(out, length_prefixed=False) == ref, "%d GPUs" % devices


def test_4k_stream_1024_frames_identical_on_1_2_4_8_gpus(oracle, b2):
    """config C4: 3840x2160, closed GOPs of 32, 1,024 frames (32 GOPs) through the x264-mirror on 1, 2, 4, 8 GPUs"""
    import time
    w, h, n, gop = 3840, 2160, 1024, 32
    counts = device_counts(b2)
    base = [oracle.synth_frame(w, h, t, 0) for t in range(32)]
    digests, rates = {}, {}
    for devices in counts:
        enc = b2.DropInEncoder(w, h, preset="slow", tune="film", quality=26, fps=(60, 1), annexb=0, i_keyint_max=gop, i_gop_slots=8,
                               i_devices=devices)
        hsh = hashlib.sha256(); nout = 0; pts_ok = True
        t0 = time.perf_counter()
        for t in range(n):
            fr = base[t % 32] if (t // 32) % 2 == 0 else base[31 - t % 32]
            size, nals, pts, dts, key = enc.encode(fr, t)
            assert size >= 0
            if size > 0:
                pts_ok &= pts == nout; nout += 1
                for _, d in nals: hsh.update(d)
        while enc.delayed() > 0:
            size, nals, pts, dts, key = enc.encode(None, 0)
            assert size > 0
            pts_ok &= pts == nout; nout += 1
            for _, d in nals: hsh.update(d)
        rates[devices] = n / (time.perf_counter() - t0)
        enc.close()
        assert nout == n and pts_ok
        digests[devices] = hsh.hexdigest()
    print("4K drop-in frames/s by GPU count (Python driver, 12 MB numpy copy per call included):", {k: round(v) for k, v in rates.items()})
    assert len(set(digests.values())) == 1, digests
