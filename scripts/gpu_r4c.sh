#!/bin/bash
# round 2, session 3: B2ENC_STATS=1 -- where the drop-in's host threads spend their time at 16 / 32 GOP slots (result-copy buffer pool in)
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
run() {
  env $1 B2ENC_STATS=1 LD_LIBRARY_PATH=video-encoder_b200 tools/b2_encode --size 1920x1080 --fps 60 --quality 26 --gop 32 $2 $IN /dev/shm/b2_probe_out.h264 > gpurun_out/cli_probe.out 2> gpurun_out/cli_probe.err || { echo "failed: $2"; tail -3 gpurun_out/cli_probe.err; }
  echo "$1 b2_encode $2 | loop: $(tail -1 gpurun_out/cli_probe.out)"; grep "b2enc stats" gpurun_out/cli_probe.err
}
{
run "WARMUP=1" "--preset slow --slots 16" | head -1
run "A=1" "--preset slow --slots 16"
run "A=1" "--preset slow --slots 32"
run "A=1" "--preset slow --slots 48"
run "A=1" "--preset slow --profile baseline --slots 32"
} | tee gpurun_out/r4c_cli_stats.txt
rm -f $IN /dev/shm/b2_probe_out.h264
