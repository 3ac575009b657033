"""throughput of the x264-mirror drop-in encoder (GPU stage + host entropy workers) at 1080p through the reference's call sequence"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, b2enc, b2oracle as o
W, H, N = 1920, 1080, int(os.environ.get("B2_PROBE_FRAMES", "2048"))
frames = [o.synth_frame(W, H, t % 16) for t in range(16)]
for name, kw in (("baseline (CAVLC)", dict(profile="baseline")), ("main (CABAC, default)", dict()),
                 ("high (CABAC + 8x8 + partitions 2)", dict(profile="high", b_transform_8x8=1, b_partitions=2))):
    for slots in (4, 8, 16):
        enc = b2enc.DropInEncoder(W, H, preset="slow", tune="film", quality=26, fps=(60, 1), annexb=0, i_keyint_max=32, i_gop_slots=slots, **kw)
        t0 = time.perf_counter(); nbytes = 0; nout = 0
        for t in range(N):
            size, nals, pts, dts, key = enc.encode(frames[t % 16], t)
            if size > 0: nbytes += size; nout += 1
        while enc.delayed() > 0:
            size, nals, pts, dts, key = enc.encode(None, 0)
            nbytes += size; nout += 1
        dt = time.perf_counter() - t0
        enc.close()
        print("%-36s slots %2d: %d frames, %.2f s = %.0f frames/s, %.1f Mbit/s at 60 fps" % (name, slots, nout, dt, nout / dt, nbytes * 8 * 60 / nout / 1e6), flush=True)
