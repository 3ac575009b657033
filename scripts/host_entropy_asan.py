"""Host slice writers under AddressSanitizer + UBSan: random oracle-encoded frames (sizes, QPs, CAVLC/CABAC, deblocking, 8x8
transform, partitions, content types) are dumped with exact-size buffers and written by scripts/host_entropy_asan.c from the dense
and the packed levels.   usage: host_entropy_asan.py [cases] [seed]   (CPU only)"""
import sys, os, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("oracle", "tests", "video-encoder_b200"):
    sys.path.insert(0, os.path.join(ROOT, d))
import numpy as np
import b2oracle as o, b2enc
from test_oracle_decode import smooth_seq, coarse_seq, shear_seq

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
work = tempfile.mkdtemp(prefix="b2_asan_")
man = []; idx = 0
for case in range(ncases):
    w = int(rng.integers(8, 70)) * 2; h = int(rng.integers(8, 50)) * 2
    qp = int(rng.integers(10, 52)); T = int(rng.integers(2, 4))
    cabac = int(rng.random() < 0.6); deblock = int(rng.random() < 0.5); t8 = int(rng.random() < 0.5); parts = int(rng.choice([0, 1, 2]))
    kind = str(rng.choice(["smooth", "coarse", "shear", "noise"])); sd = int(rng.integers(0, 1 << 30))
    if kind == "smooth": fr = smooth_seq(w, h, T, seed=sd, cut=1 if rng.random() < 0.4 else None)
    elif kind == "coarse": fr = coarse_seq(w, h, T, seed=sd, scale=int(rng.integers(4, 20)))
    elif kind == "shear": fr = shear_seq(w, h, T, seed=sd, amp=int(rng.integers(1, 3)))
    else:
        r = np.random.default_rng(sd)
        fr = [(r.integers(0, 256, (h, w), dtype=np.uint8), r.integers(0, 256, (h // 2, w // 2), dtype=np.uint8), r.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)) for _ in range(T)]
    bs, recons, infos, coefs = o.encode_sequence(fr, w, h, qp=qp, merange=16, gop=32, deblock=deblock, cabac=cabac, transform8x8=t8, partitions=parts)
    mbw, mbh = (w + 15) // 16, (h + 15) // 16
    for t, (info, coef) in enumerate(zip(infos, coefs)):
        info.tofile(os.path.join(work, f"info{idx}.bin")); coef.tofile(os.path.join(work, f"coef{idx}.bin"))
        p = b2enc.pack_levels(info, coef); p.tofile(os.path.join(work, f"packed{idx}.bin"))
        man.append(f"{idx} {w} {h} {mbw} {mbh} {qp} {0 if t == 0 else 1} {t} {cabac} {deblock} {t8} {p.size}")
        idx += 1
open(os.path.join(work, "manifest.txt"), "w").write("\n".join(man) + "\n")
host = os.path.join(ROOT, "video-encoder_b200", "host")
exe = os.path.join(work, "driver")
subprocess.check_call(["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I" + os.path.join(ROOT, "include"), "-I" + host,
                       "-o", exe, os.path.join(ROOT, "scripts", "host_entropy_asan.c"), os.path.join(host, "b2h_cavlc.c"),
                       os.path.join(host, "b2h_cabac.c"), os.path.join(host, "b2h_avcc.c"), "-lpthread"])
r = subprocess.run([exe, work], capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1"))
print(r.stdout.strip()); 
if r.returncode != 0 or "no sanitizer finding" not in r.stdout:
    print(r.stderr[-3000:]); sys.exit(1)
import shutil; shutil.rmtree(work, ignore_errors=True)
