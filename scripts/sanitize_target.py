"""small end-to-end passes for compute-sanitizer: named path and all features (deblocking, 8x8 transform + intra 8x8, partitions,
packed levels), odd sizes, every raw input format"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, b2enc, b2oracle
from test_oracle_decode import smooth_seq, coarse_seq, shear_seq
for (w, h, R, full, qp, gen) in [(144, 112, 16, 0, 30, smooth_seq), (150, 98, 32, 1, 30, shear_seq), (96, 80, 16, 1, 42, coarse_seq)]:
    seqs = [gen(w, h, 3, seed=s) for s in range(2)]
    seqs[1][1] = smooth_seq(w, h, 1, seed=9)[0]                     # scene change -> intra MBs in a P frame
    eng = b2enc.Engine(w, h, slots=2, ring=1, merange=R, qp=qp, subpel=1, intra_in_p=1, deblock=1, transform8x8=full, partitions=full,
                       pack_levels=full)
    for t in range(3):
        for s in range(2): eng.put_frame(s, 0, list(seqs[s][t]))
        eng.h2d(); eng.encode(b2enc.FRAME_I if t == 0 else b2enc.FRAME_P); eng.d2h(); eng.sync()
        if full: eng.packed(0)
    eng.close()
rng = np.random.default_rng(0)
b2enc.sws_convert("yuyv422", 34, 18, [rng.integers(0, 256, (18, 68), dtype=np.uint8)])
b2enc.sws_convert("bgr24", 34, 19, [rng.integers(0, 256, (19, 102), dtype=np.uint8)])
b2enc.sws_convert("rgb24", 34, 18, [rng.integers(0, 256, (18, 102), dtype=np.uint8)])
b2enc.sws_convert("yuv422p", 35, 18, [rng.integers(0, 256, (18, 35), dtype=np.uint8), rng.integers(0, 256, (18, 18), dtype=np.uint8), rng.integers(0, 256, (18, 18), dtype=np.uint8)])
b2enc.sws_convert("yuv411p", 36, 18, [rng.integers(0, 256, (18, 36), dtype=np.uint8), rng.integers(0, 256, (18, 9), dtype=np.uint8), rng.integers(0, 256, (18, 9), dtype=np.uint8)])
print("sanitize target ok")
