"""time the K1 variants built by scripts/k1_variants.sh (1080p, +-32, 32 frames per launch like bench.py); each variant in its own process"""
import os, subprocess, sys, glob, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np, b2enc, b2oracle, hashlib
    w, h, R, n = 1920, 1088, 32, 32
    cur = np.stack([b2oracle.synth_frame(w, h, t + 1, t % 5)[0] for t in range(n)])
    ref = np.stack([b2oracle.synth_frame(w, h, t, t % 5)[0] for t in range(n)])
    mv, cost, ms = b2enc.me_fullpel(cur, ref, R, lam=4, iters=10)
    peak, _ = b2enc.vabsdiff4_peak(0, 512, 5)
    work = n * (w // 16) * (h // 16) * 65 * 65 * 256
    print(json.dumps({"ms": ms, "frac": work / (ms * 1e-3) / (peak * 4), "hash": hashlib.md5(mv.tobytes() + cost.tobytes()).hexdigest()[:8]}))
    sys.exit(0)
res = []
for lib in sorted(glob.glob(os.path.join(ROOT, "build_variants", "*.so"))):
    for nt in (256,):
        env = dict(os.environ, B2ENC_LIB=lib, B2_K1_THREADS=str(nt))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip()[-200:]
        print(os.path.basename(lib), nt, line, flush=True)
