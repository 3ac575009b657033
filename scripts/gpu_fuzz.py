"""Randomised parity sweep on the GPU: engine (C-ABI) vs C oracle over random picture sizes, QPs, search ranges, tool
sets and content types, frame by frame and field by field (tests/test_engine_parity.run_and_compare).
usage: gpu_fuzz.py [cases] [seed]   -- prints one line per case and a summary; exits 1 on the first mismatch"""
import sys, os, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("video-encoder_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, d))
import numpy as np
import b2enc, b2oracle
from test_engine_parity import run_and_compare
from test_oracle_decode import smooth_seq, coarse_seq, shear_seq

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed0)
b2oracle.lib(); b2enc.lib()


def noise_seq(w, h, n, seed):
    r = np.random.default_rng(seed)
    out = []
    for t in range(n):
        y = r.integers(0, 256, (h, w), dtype=np.uint8)
        out.append((y, r.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8), r.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8)))
    return out


def flat_seq(w, h, n, seed):
    r = np.random.default_rng(seed)
    vals = r.integers(0, 256, (n, 3))
    return [(np.full((h, w), v[0], np.uint8), np.full(((h + 1) // 2, (w + 1) // 2), v[1], np.uint8), np.full(((h + 1) // 2, (w + 1) // 2), v[2], np.uint8)) for v in vals]


t0 = time.time(); done = 0
for case in range(ncases):
    w = int(rng.integers(8, 100)) * 4 if rng.random() < 0.7 else int(rng.integers(16, 200)) * 2
    h = int(rng.integers(8, 80)) * 4 if rng.random() < 0.7 else int(rng.integers(16, 160)) * 2
    if rng.random() < 0.15: w = int(rng.integers(8, 33)) * 2             # tiny pictures: 1-4 macroblocks per row / column
    if rng.random() < 0.15: h = int(rng.integers(8, 33)) * 2
    if rng.random() < 0.12: h = int(rng.integers(80, 330)) * 2; w = min(w, 128)   # tall: K8's row pipeline spans several CTAs of a cluster
    qp = int(rng.integers(10, 52)); R = int(rng.choice([16, 32]))
    subpel = int(rng.random() < 0.85); intra = int(rng.random() < 0.85)
    deblock = int(rng.random() < 0.5); t8 = int(rng.random() < 0.5); parts = int(rng.choice([0, 1, 2])); pack = int(rng.random() < 0.5)
    if not subpel: parts = 0
    offs = (int(rng.integers(-6, 7)), int(rng.integers(-6, 7))) if deblock and rng.random() < 0.6 else (0, 0)     # loop-filter offsets (tunes)
    kind = str(rng.choice(["smooth", "coarse", "shear", "noise", "flat", "synth"]))
    S = int(rng.integers(1, 4)); T = int(rng.integers(2, 5)); sd = int(rng.integers(0, 1 << 30))
    seqs = []
    for s in range(S):
        if kind == "smooth": seqs.append(smooth_seq(w, h, T, seed=sd + s, cut=(int(rng.integers(1, T)) if rng.random() < 0.3 else None)))
        elif kind == "coarse": seqs.append(coarse_seq(w, h, T, seed=sd + s, scale=int(rng.integers(4, 24))))
        elif kind == "shear": seqs.append(shear_seq(w, h, T, seed=sd + s, stripe=int(rng.integers(12, 60)), band=int(rng.integers(12, 60)), amp=int(rng.integers(1, 4))))
        elif kind == "noise": seqs.append(noise_seq(w, h, T, sd + s))
        elif kind == "flat": seqs.append(flat_seq(w, h, T, sd + s))
        else: seqs.append([b2oracle.synth_frame(w, h, t, s + sd % 7) for t in range(T)])
    prune = int(os.environ.get("B2_FUZZ_PRUNE", "-1"))                   # -1: random; 1: always (with +-32: the only range the engine prunes at)
    if prune < 0: prune = int(rng.random() < 0.5)
    elif prune: R = 32; parts = min(parts, 1)
    desc = f"case {case}: {w}x{h} qp {qp} R {R} subpel {subpel} intra {intra} deblock {deblock}{offs} t8 {t8} parts {parts} pack {pack} prune {prune} {kind} S {S} T {T} seed {sd}"
    try:
        run_and_compare(b2oracle, b2enc, seqs, w, h, qp, R, subpel=subpel, intra_in_p=intra, deblock=deblock, transform8x8=t8,
                        pack_levels=pack, partitions=parts, deblock_offsets=offs, me_prune=prune)
    except Exception:
        print("MISMATCH/ERROR", desc, flush=True); traceback.print_exc(); sys.exit(1)
    done += 1
    print("ok", desc, flush=True)
print(f"{done} cases bit-exact in {time.time() - t0:.0f} s")
