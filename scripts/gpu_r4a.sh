#!/bin/bash
# round 2, session 3: the reference's call sequence through tools/b2_encode with the faster host CABAC writer, GOP-slot sweep
mkdir -p gpurun_out
IN=/dev/shm/b2_probe_1080p.yuv
N=3072
python - "$IN" "$N" <<'PY'
import sys, os
ROOT = os.getcwd()
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b2oracle as o
fr = [b"".join(p.tobytes() for p in o.synth_frame(1920, 1080, t)) for t in range(32)]
with open(sys.argv[1], "wb") as f:
    for i in range(int(sys.argv[2])): f.write(fr[i % 32] if (i // 32) % 2 == 0 else fr[31 - i % 32])
PY
nproc
run() {
  s=$(date +%s.%N)
  env $1 LD_LIBRARY_PATH=video-encoder_b200 tools/b2_encode --size 1920x1080 --fps 60 --quality 26 --gop 32 $2 $IN /dev/shm/b2_probe_out.h264 > gpurun_out/cli_probe.out 2> gpurun_out/cli_probe.err || { echo "failed: $2"; tail -3 gpurun_out/cli_probe.err; }
  e=$(date +%s.%N)
  python -c "dt=$e-$s; print('%-18s b2_encode %-46s process %.2f s | loop: %s | sha %s' % ('$1', '$2', dt, open('gpurun_out/cli_probe.out').read().strip().splitlines()[-1], __import__('hashlib').sha256(open('/dev/shm/b2_probe_out.h264','rb').read()).hexdigest()[:12]))"
}
{
run "A=1" "--preset slow --slots 16"
run "A=1" "--preset slow --slots 24"
run "A=1" "--preset slow --slots 32"
run "B2ENC_ME_PRUNE=1" "--preset slow --slots 16"
run "B2ENC_ME_PRUNE=1" "--preset slow --slots 32"
run "A=1" "--preset slow --8x8dct --partitions 2 --slots 32"
run "A=1" "--preset slow --profile baseline --slots 32"
} | tee gpurun_out/r4a_cli.txt
rm -f $IN /dev/shm/b2_probe_out.h264
