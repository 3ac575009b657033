"""time K1 at +-16 for the strip widths built into build_variants/ (B2_K1_NMB16) and CTA sizes (B2_K1_THREADS): 720p and 1080p luma, 32 frames
per launch; each variant in its own process.  Bit-exactness of every variant is checked through the result hash."""
import os, subprocess, sys, glob, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np, b2enc, b2oracle, hashlib
    out = {}
    peak, _ = b2enc.vabsdiff4_peak(0, 512, 5)
    for (w, h) in ((1280, 720), (1920, 1088)):
        n, R = 32, 16
        cur = np.stack([b2oracle.synth_frame(w, h, t + 1, t % 5)[0] for t in range(n)])
        ref = np.stack([b2oracle.synth_frame(w, h, t, t % 5)[0] for t in range(n)])
        mv, cost, ms = b2enc.me_fullpel(cur, ref, R, lam=4, iters=10)
        work = n * (w // 16) * (h // 16) * 33 * 33 * 256
        out["%dx%d" % (w, h)] = {"ms": round(ms, 4), "frac": round(work / (ms * 1e-3) / (peak * 4), 4), "hash": hashlib.md5(mv.tobytes() + cost.tobytes()).hexdigest()[:8]}
    print(json.dumps(out))
    sys.exit(0)
for lib in sorted(glob.glob(os.path.join(ROOT, "build_variants", "libb2enc_nmb16_*.so"))):
    for nt in (192, 256, 320, 384):
        env = dict(os.environ, B2ENC_LIB=lib, B2_K1_THREADS=str(nt))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip()[-200:]
        print(os.path.basename(lib), nt, line, flush=True)
