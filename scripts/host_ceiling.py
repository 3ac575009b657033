"""Host-stage ceiling of the drop-in encoder, CPU only: the product's host sources (video-encoder_b200/host/*.c) linked with the
zero-latency mock engine (tests/mock/mock_engine.c, B2_MOCK_CANNED) and driven by scripts/host_ceiling.c with the reference's
per-frame pair sws_scale + encoder_encode.  The canned results are oracle encodes of the bench's synthetic content (1 I + 3 P
frames, QP 26) at the given size, so the entropy workers code realistic slices.  What it measures: the frame rate the host
stage could carry if the GPUs took no time -- the ceiling of one stream on this host, whatever the number of GPUs.
B2_CEILING_CONTENT=static: a still picture (P frames are all skip), so that the entropy stage costs next to nothing and the
caller's thread / the per-GPU threads show their own limits.
usage: host_ceiling.py [WIDTH HEIGHT [FRAMES [DEVICES [SLOTS]]]]   (B2ENC_STATS=1 etc. are passed through)"""
import sys, os, glob, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("oracle", "video-encoder_b200"):
    sys.path.insert(0, os.path.join(ROOT, d))
import b2oracle as o, b2enc

W = int(sys.argv[1]) if len(sys.argv) > 2 else 1920
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
frames = sys.argv[3] if len(sys.argv) > 3 else "2048"
rest = sys.argv[4:6]
work = os.environ.get("B2_CEILING_DIR") or tempfile.mkdtemp(prefix="b2_ceiling_")
static = os.environ.get("B2_CEILING_CONTENT") == "static"
canned = os.path.join(work, "canned_%dx%d%s" % (W, H, "_static" if static else ""))
if not os.path.exists(os.path.join(canned, "packed3.bin")):
    os.makedirs(canned, exist_ok=True)
    fr = [o.synth_frame(W, H, 0 if static else t, 0) for t in range(4)]
    # the mirror's defaults at preset slow / tune film: CABAC, loop filter on (-1:-1), +-32 (the search range changes little here: +-16 is quicker)
    _, _, infos, coefs = o.encode_sequence(fr, W, H, qp=26, merange=16, gop=32, deblock=1, cabac=1)
    for t, (info, coef) in enumerate(zip(infos, coefs)):
        info.tofile(os.path.join(canned, "info%d.bin" % t)); b2enc.pack_levels(info, coef).tofile(os.path.join(canned, "packed%d.bin" % t))
so = os.path.join(work, "libb2enc_null.so"); exe = os.path.join(work, "host_ceiling")
host = os.path.join(ROOT, "video-encoder_b200", "host")
srcs = [os.path.join(ROOT, "tests", "mock", "mock_engine.c")] + sorted(glob.glob(os.path.join(host, "*.c"))) + sorted(glob.glob(os.path.join(ROOT, "oracle", "b2o_*.c")))
inc = ["-I" + os.path.join(ROOT, "include"), "-I" + host]
subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-Wall"] + inc + ["-o", so] + srcs + ["-lm", "-lpthread"])
subprocess.check_call(["gcc", "-O2", "-std=c99", "-Wall"] + inc + ["-o", exe, os.path.join(ROOT, "scripts", "host_ceiling.c"), so, "-Wl,-rpath," + work, "-lpthread"])
env = dict(os.environ, B2_MOCK_CANNED=canned, B2_MOCK_DEVICES=rest[0] if rest else "1")
sys.exit(subprocess.call([exe, str(W), str(H), frames] + rest, env=env))
