import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, b2enc, b2oracle
w, h, R, ring, steps, GOP = 1920, 1080, 32, 4, 32, 32
frames = {}
for (S, G) in [(16, 4), (16, 8), (32, 4), (32, 8), (48, 8), (64, 8)]:
    eng = b2enc.Engine(w, h, slots=S, ring=ring, merange=R, qp=26, subpel=1, intra_in_p=1, profile=0, streams=G)
    for s in range(S):
        for r in range(ring):
            if (s, r) not in frames: frames[(s, r)] = b2oracle.synth_frame(w, h, r, s)
            y, u, v = frames[(s, r)]
            buf = eng.host_input(s, r)
            buf[:w * h] = y.ravel(); buf[w * h:w * h + u.size] = u.ravel(); buf[w * h + u.size:] = v.ravel()
    for r in range(ring): eng.h2d(ring=r)
    eng.sync()
    NG = len(eng.groups()); phase = [g * GOP // NG for g in range(NG)]
    def issue(step):
        for g in range(NG):
            ft = b2enc.FRAME_I if step == 0 or (step + phase[g]) % GOP == 0 else b2enc.FRAME_P
            eng.encode_group(g, ft, ring=step % ring)
    st = 0
    for _ in range(3): issue(st); st += 1
    eng.sync(); eng.timer_start()
    for _ in range(steps): issue(st); st += 1
    ms = eng.timer_stop()
    print(json.dumps({"slots": S, "groups": NG, "fps": S * steps / ms * 1e3, "ms_per_step": ms / steps}))
    eng.close()
