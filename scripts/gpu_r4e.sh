#!/bin/bash
# round 2, session 3: entropy workers at nice 10 vs 0, staging helpers' polling time, fewer workers (CLI with B2ENC_STATS)
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
  echo "$1 b2_encode $2 | loop: $(tail -1 gpurun_out/cli_probe.out) | sha $(sha256sum /dev/shm/b2_probe_out.h264 | cut -c1-12)"; grep "b2enc stats" gpurun_out/cli_probe.err
}
{
run "WARMUP=1" "--preset slow" | head -1
run "A=1" "--preset slow"
run "B2ENC_WORKER_NICE=0" "--preset slow"
run "B2ENC_SWS_SPIN=8000" "--preset slow"
run "B2ENC_SWS_SPIN=0" "--preset slow"
run "B2ENC_ENTROPY_THREADS=11" "--preset slow"
run "A=1" "--preset slow"
} | tee gpurun_out/r4e_cli_stats.txt
rm -f $IN /dev/shm/b2_probe_out.h264
