"""throughput of the pre-filter stage on DV-PAL sized pictures (720x576 yuv420p) through the C-ABI (host frames in / out)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, b2enc
from test_oracle_decode import smooth_seq
w, h, n = 720, 576, 100
frames = smooth_seq(w, h, 4, seed=1)
for spec in ("hqdn3d", "yadif", "hqdn3d,yadif"):
    g = b2enc.FilterGraph(w, h, spec)
    g.run(frames)                      # warm-up
    t0 = time.perf_counter()
    got = 0
    for t in range(n):
        g.add(frames[t % 4], pts=t)
        while g.poll() > 0:
            g.get(); got += 1
    g.flush()
    while g.poll() > 0:
        g.get(); got += 1
    dt = time.perf_counter() - t0
    g.close()
    print("%-14s %d frames in %.3f s = %.0f frames/s (push + pull through host memory, one stream)" % (spec, got, dt, got / dt))
