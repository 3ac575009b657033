import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, b2enc, b2oracle
rate, ms = b2enc.vabsdiff4_peak(0, 512, 5)
print(json.dumps({"vabsdiff4_lane_instr_per_s": rate, "pix_sad_per_s": rate * 4, "ms": ms}))
for (w, h, R, n) in [(1280, 720, 16, 8), (1920, 1088, 32, 4), (1920, 1088, 32, 16)]:
    h16 = (h + 15) // 16 * 16
    cur = np.stack([b2oracle.synth_frame(w, h16, t + 1)[0] for t in range(n)])
    ref = np.stack([b2oracle.synth_frame(w, h16, t)[0] for t in range(n)])
    mv, cost, kms = b2enc.me_fullpel(cur, ref, R, iters=10)
    sads = (w // 16) * (h16 // 16) * (2 * R + 1) ** 2 * 256 * n
    print(json.dumps({"w": w, "h": h16, "R": R, "frames": n, "kernel_ms": kms, "fps": n / kms * 1e3,
                      "Tpix_sad_per_s": sads / kms / 1e9, "frac_of_peak": sads / (kms * 1e-3) / (rate * 4)}))
