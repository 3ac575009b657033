"""one 1080p slot, deblocking on: I, P, P -- target for ncu captures of the K7 / K8 / K9 latency-chain kernels"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, b2enc, b2oracle
w, h, S = 1920, 1080, 1
eng = b2enc.Engine(w, h, slots=S, ring=3, merange=32, qp=26, subpel=1, intra_in_p=1, streams=1, pack_levels=1, deblock=1, deblock_offsets=(-1, -1))
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
