"""Randomised decoder round trip on the CPU: oracle encode stage + host slice writers (CAVLC / CABAC, all tool sets) ->
libavcodec's H.264 decoder -> must equal the oracle's reconstruction bit for bit.  Pins the normative arithmetic and the
entropy writers on random sizes / QPs / content.   usage: cpu_fuzz_entropy.py [cases] [seed]"""
import sys, os, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, d))
import numpy as np
import b2oracle
from test_oracle_decode import smooth_seq, coarse_seq, shear_seq, _roundtrip

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 50
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
for case in range(ncases):
    w = int(rng.integers(8, 80)) * 2; h = int(rng.integers(8, 64)) * 2
    qp = int(rng.integers(10, 52)); R = int(rng.choice([16, 32])); T = int(rng.integers(2, 6))
    cabac = int(rng.random() < 0.6); deblock = int(rng.random() < 0.5); t8 = int(rng.random() < 0.5); parts = int(rng.choice([0, 1, 2]))
    subpel = int(rng.random() < 0.85)
    if not subpel: parts = 0
    kind = str(rng.choice(["smooth", "coarse", "shear", "noise", "synth"])); sd = int(rng.integers(0, 1 << 30))
    if kind == "smooth": fr = smooth_seq(w, h, T, seed=sd, cut=(int(rng.integers(1, T)) if rng.random() < 0.4 else None))
    elif kind == "coarse": fr = coarse_seq(w, h, T, seed=sd, scale=int(rng.integers(4, 24)))
    elif kind == "shear": fr = shear_seq(w, h, T, seed=sd, stripe=int(rng.integers(12, 60)), band=int(rng.integers(12, 60)), amp=int(rng.integers(1, 4)))
    elif kind == "noise":
        r = np.random.default_rng(sd)
        fr = [(r.integers(0, 256, (h, w), dtype=np.uint8), r.integers(0, 256, (h // 2, w // 2), dtype=np.uint8), r.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)) for _ in range(T)]
    else: fr = [b2oracle.synth_frame(w, h, t, sd % 7) for t in range(T)]
    desc = f"case {case}: {w}x{h} qp {qp} R {R} subpel {subpel} cabac {cabac} deblock {deblock} t8 {t8} parts {parts} {kind} T {T} seed {sd}"
    try:
        _roundtrip(b2oracle, fr, w, h, qp=qp, merange=R, subpel=subpel, gop=int(rng.choice([2, 3, 32])), deblock=deblock, cabac=cabac,
                   transform8x8=t8, partitions=parts)
    except Exception:
        print("MISMATCH/ERROR", desc, flush=True); traceback.print_exc(); sys.exit(1)
    print("ok", desc, flush=True)
print(f"{ncases} cases decoder-exact in {time.time() - t0:.0f} s")
