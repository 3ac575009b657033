import sys, os
sys.path.insert(0, "video-encoder_b200"); sys.path.insert(0, "oracle")
import numpy as np, b2enc, b2oracle
w, h, R, n = 1920, 1088, 32, 16
cur = np.stack([b2oracle.synth_frame(w, h, t + 1)[0] for t in range(n)])
ref = np.stack([b2oracle.synth_frame(w, h, t)[0] for t in range(n)])
for lam in (0, -12345):
    mv, cost, kms = b2enc.me_fullpel(cur, ref, R, lam=lam, iters=10)
    print("lambda", lam, "kernel_ms", kms)
