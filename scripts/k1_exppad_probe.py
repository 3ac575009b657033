"""A/B of the expanded-window pitch padding in K1 (B2_K1_EXP_PAD, build_variants/libb2enc_exppad{0,1}.so): +-32 at 1080p / 4K and
+-16 at 720p / 1080p, 32 (8 at 4K) frames per launch, each variant in its own process; the result hash shows both are bit-identical."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np, b2enc, b2oracle, hashlib
    out = {}
    peak, _ = b2enc.vabsdiff4_peak(0, 512, 5)
    for (w, h, R, n) in ((1920, 1088, 32, 32), (3840, 2160, 32, 8), (1280, 720, 16, 32), (1920, 1088, 16, 32)):
        cur = np.stack([b2oracle.synth_frame(w, h, t + 1, t % 5)[0] for t in range(n)])
        ref = np.stack([b2oracle.synth_frame(w, h, t, t % 5)[0] for t in range(n)])
        mv, cost, ms = b2enc.me_fullpel(cur, ref, R, lam=4, iters=10)
        nd = 2 * R + 1
        work = n * (w // 16) * (h // 16) * nd * nd * 256
        out["%dx%d+-%d" % (w, h, R)] = {"ms": round(ms, 4), "frac": round(work / (ms * 1e-3) / (peak * 4), 4), "hash": hashlib.md5(mv.tobytes() + cost.tobytes()).hexdigest()[:8]}
    print(json.dumps(out))
    sys.exit(0)
for rep in range(2):
    for pad in (0, 1):
        lib = os.path.join(ROOT, "build_variants", "libb2enc_exppad%d.so" % pad)
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, B2ENC_LIB=lib), capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip()[-300:]
        print("pad", pad, line, flush=True)
