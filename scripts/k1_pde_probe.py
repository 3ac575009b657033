"""K1 (+K1a) time of the pruned search inside the engine -- reference = reconstruction at QP 26, predictors = previous vectors, the
bench's picture ring -- for every build in build_variants/ (partial-distortion schedules etc.); one child process per library.
1080p, +-32, 16 slots on one stream, 1 I + 7 P steps of the 8-picture ring timed after one warm-up pass."""
import os, sys, glob, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np, b2enc, b2oracle
    w, h, S, RING = 1920, 1080, 16, 8
    eng = b2enc.Engine(w, h, slots=S, ring=RING, merange=32, qp=26, subpel=1, intra_in_p=1, streams=1, pack_levels=1, profile=1, me_prune=1)
    for s in range(S):
        for r in range(RING):
            y, u, v = b2oracle.synth_frame(w, h, r, s)
            buf = eng.host_input(s, r)
            buf[:w * h] = y.ravel(); buf[w * h:w * h + u.size] = u.ravel(); buf[w * h + u.size:] = v.ravel()
    for r in range(RING): eng.h2d(ring=r)
    eng.encode(b2enc.FRAME_I, ring=0)
    for i in range(1, 8): eng.encode(b2enc.FRAME_P, ring=i)
    eng.sync(); eng.profile_reset()
    a0, b0 = eng.k1_stats()
    for i in range(8, 24): eng.encode(b2enc.FRAME_P, ring=i % RING)        # two rounds of the ring: 2 of 16 steps wrap (wrong predictors)
    ms = eng.kernel_ms()["K1 full-pel SAD"]
    a1, b1 = eng.k1_stats()
    print(json.dumps({"k1_ms_per_16_frame_step": round(ms[0] / ms[1], 4), "executed_fraction": round((a1 - a0) / (b1 - b0), 4)}))
    eng.close()
    sys.exit(0)
for lib in [None] + sorted(glob.glob(os.path.join(ROOT, "build_variants", "libb2enc_*.so"))):
    env = dict(os.environ)
    if lib: env["B2ENC_LIB"] = lib
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip()[-300:]
    print(os.path.basename(lib) if lib else "libb2enc.so (default)", line, flush=True)
