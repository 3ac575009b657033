"""small end-to-end pass (I + 2 P, two slots, deblocking on, odd size) for compute-sanitizer"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, b2enc, b2oracle
from test_oracle_decode import smooth_seq
for (w, h, R) in [(144, 112, 16), (150, 98, 32)]:
    seqs = [smooth_seq(w, h, 3, seed=s, cut=(1 if s else None)) for s in range(2)]
    eng = b2enc.Engine(w, h, slots=2, ring=1, merange=R, qp=30, subpel=1, intra_in_p=1, deblock=1)
    for t in range(3):
        for s in range(2): eng.put_frame(s, 0, list(seqs[s][t]))
        eng.h2d(); eng.encode(b2enc.FRAME_I if t == 0 else b2enc.FRAME_P); eng.d2h(); eng.sync()
    eng.close()
y, u, v = b2enc.sws_convert("yuyv422", 34, 18, [np.random.default_rng(0).integers(0, 256, (18, 68), dtype=np.uint8)])
print("sanitize target ok")
