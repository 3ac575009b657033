"""Per-slice time of the host entropy writers on the CPU (profiles/r2s3_host_stage.md): oracle-encodes a few frames of the bench's
synthetic content (1 I + 3 P) at WIDTH x HEIGHT, QP 26 -- once with the mirror's default tools, once with the 8x8 transform and
partitions, once as a still picture (all-skip P slices) -- and times scripts/entropy_timing.c on them, CABAC and CAVLC.
usage: entropy_timing.py [WIDTH HEIGHT [REPS]]      (pin the process, e.g. taskset -c 3, on a noisy host)"""
import sys, os, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("oracle", "video-encoder_b200"):
    sys.path.insert(0, os.path.join(ROOT, d))
import b2oracle as o, b2enc
W = int(sys.argv[1]) if len(sys.argv) > 2 else 1920
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
reps = sys.argv[3] if len(sys.argv) > 3 else "30"
work = tempfile.mkdtemp(prefix="b2_entropy_")
mbw, mbh = (W + 15) // 16, (H + 15) // 16
cases = (("default tools", dict(transform8x8=0, partitions=0), False), ("8x8 transform + partitions", dict(transform8x8=1, partitions=1), False),
         ("still picture", dict(transform8x8=0, partitions=0), True))
host = os.path.join(ROOT, "video-encoder_b200", "host")
exe = os.path.join(work, "entropy_timing")
subprocess.check_call(["gcc", "-O2", "-std=c99", "-I" + os.path.join(ROOT, "include"), "-I" + host, "-o", exe, os.path.join(ROOT, "scripts", "entropy_timing.c"),
                       os.path.join(host, "b2h_cavlc.c"), os.path.join(host, "b2h_cabac.c"), os.path.join(host, "b2h_avcc.c"), "-lpthread"])
for k, (name, kw, still) in enumerate(cases):
    fr = [o.synth_frame(W, H, 0 if still else t, 0) for t in range(4)]
    _, _, infos, coefs = o.encode_sequence(fr, W, H, qp=26, merange=16, gop=32, deblock=1, cabac=1, **kw)
    d = os.path.join(work, "case%d" % k); os.makedirs(d)
    man = []
    for t, (info, coef) in enumerate(zip(infos, coefs)):
        info.tofile(os.path.join(d, "info%d.bin" % t)); p = b2enc.pack_levels(info, coef); p.tofile(os.path.join(d, "packed%d.bin" % t))
        man.append("%d %d %d %d %d 26 %d %d 1 1 %d %d" % (t, W, H, mbw, mbh, 0 if t == 0 else 1, t, kw["transform8x8"], p.size))
    open(os.path.join(d, "manifest.txt"), "w").write("\n".join(man) + "\n")
    for cabac in (1, 0):
        print("--- %s, %s" % (name, "CABAC" if cabac else "CAVLC"), flush=True)
        subprocess.check_call([exe, d, reps, str(cabac)])
