"""K1 vs its partition variant at 1080p +-32: kernel time of one 16-frame launch"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, b2enc, b2oracle
w, h, R, n = 1920, 1088, 32, 16
cur = np.stack([b2oracle.synth_frame(w, h, t + 1)[0] for t in range(n)])
ref = np.stack([b2oracle.synth_frame(w, h, t)[0] for t in range(n)])
_, _, ms0 = b2enc.me_fullpel(cur, ref, R, lam=4, iters=10)
_, _, ms1 = b2enc.me_fullpel_parts(cur, ref, R, lam=4, iters=10)
work = n * (w // 16) * (h // 16) * 65 * 65 * 256
print("K1 16x16: %.3f ms (%.1f Tpix-SAD/s)   K1 parts: %.3f ms (%.1f Tpix-SAD/s)  ratio %.3f" % (ms0, work / ms0 / 1e9, ms1, work / ms1 / 1e9, ms1 / ms0))
