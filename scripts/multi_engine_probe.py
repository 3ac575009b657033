import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, b2enc, b2oracle
w, h, R, ring, steps = 1920, 1080, 32, 8, 32
for (E, S) in [(1, 16), (2, 8), (4, 4), (2, 16), (4, 8)]:
    engs = [b2enc.Engine(w, h, slots=S, ring=ring, merange=R, qp=26, subpel=1, intra_in_p=1, profile=0) for _ in range(E)]
    for ei, eng in enumerate(engs):
        for s in range(S):
            for r in range(ring):
                y, u, v = b2oracle.synth_frame(w, h, r, ei * S + s)
                buf = eng.host_input(s, r)
                buf[:w * h] = y.ravel(); buf[w * h:w * h + u.size] = u.ravel(); buf[w * h + u.size:] = v.ravel()
        for r in range(ring): eng.h2d(ring=r)
        eng.sync()
    def ft(i): return b2enc.FRAME_I if i % 32 == 0 else b2enc.FRAME_P
    for i in range(3):
        for eng in engs: eng.encode(ft(i), ring=i % ring)
    for eng in engs: eng.sync()
    t0 = time.perf_counter()
    for i in range(steps):
        for eng in engs: eng.encode(ft(3 + i), ring=(3 + i) % ring)
    for eng in engs: eng.sync()
    dt = time.perf_counter() - t0
    print(json.dumps({"engines": E, "slots_each": S, "fps": E * S * steps / dt, "ms_per_round": dt / steps * 1e3}))
    for eng in engs: eng.close()
