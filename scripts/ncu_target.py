"""small deterministic workload for ncu: 1080p, +-32, 4 slots on one stream: 1 I step + 2 P steps"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, b2enc, b2oracle
w, h, S = 1920, 1080, 4
FULL = int(os.environ.get("B2_ALL_FEATURES", "0"))      # 1: deblocking + 8x8 transform / intra 8x8 + partitions (rows N1, N2)
PRUNE = int(os.environ.get("B2_ME_PRUNE", "0"))         # 1: the pruned full-pel search (K1a + k1_me_fullpel_sea_kernel)
eng = b2enc.Engine(w, h, slots=S, ring=3, merange=32, qp=26, subpel=1, intra_in_p=1, streams=1, pack_levels=1,
                   deblock=FULL, transform8x8=FULL, partitions=FULL, me_prune=PRUNE)
for s in range(S):
    for r in range(3):
        y, u, v = b2oracle.synth_frame(w, h, r, s)
        buf = eng.host_input(s, r)
        buf[:w * h] = y.ravel(); buf[w * h:w * h + u.size] = u.ravel(); buf[w * h + u.size:] = v.ravel()
for r in range(3): eng.h2d(ring=r)
eng.encode(b2enc.FRAME_I, ring=0)
eng.encode(b2enc.FRAME_P, ring=1)
eng.encode(b2enc.FRAME_P, ring=2)
eng.sync()
print("ok", eng.launch_count())
eng.close()
