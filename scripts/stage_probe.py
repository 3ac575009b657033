import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, b2enc, b2oracle
w, h, R, S, ring, steps = [int(a) for a in (sys.argv[1:7] if len(sys.argv) > 6 else "1920 1080 32 16 4 8".split())]
streams = int(sys.argv[7]) if len(sys.argv) > 7 else 1
eng = b2enc.Engine(w, h, slots=S, ring=ring, merange=R, qp=26, subpel=1, intra_in_p=1, profile=1, streams=streams, deblock=int(os.environ.get('B2_DEBLOCK', '0')))
t0 = time.time()
for s in range(S):
    for r in range(ring):
        y, u, v = b2oracle.synth_frame(w, h, r, s)
        buf = eng.host_input(s, r)
        buf[:w * h] = y.ravel(); buf[w * h:w * h + u.size] = u.ravel(); buf[w * h + u.size:] = v.ravel()
print("synth %.1fs" % (time.time() - t0))
for r in range(ring):
    eng.h2d(ring=r)
eng.encode(b2enc.FRAME_I, ring=0); eng.sync()
for i in range(3):
    eng.encode(b2enc.FRAME_P, ring=(i + 1) % ring)
eng.sync(); eng.profile_reset()
eng.timer_start()
for i in range(steps):
    eng.encode(b2enc.FRAME_P, ring=i % ring)
ms = eng.timer_stop()
print(json.dumps({"w": w, "h": h, "R": R, "slots": S, "steps": steps, "ms_per_step": ms / steps, "fps": S * steps / ms * 1e3}))
for k, (kms, n) in eng.kernel_ms().items():
    if n: print("%-24s %8.3f ms/step  (%d launches)" % (k, kms / steps, n))
eng.profile_reset()
eng.timer_start(); eng.encode(b2enc.FRAME_I, ring=0); ms = eng.timer_stop()
print("I-frame step: %.3f ms" % ms)
for k, (kms, n) in eng.kernel_ms().items():
    if n: print("%-24s %8.3f ms" % (k, kms))
