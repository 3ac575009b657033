#!/bin/bash
# round 2, session 3: what bounds the reference's call sequence at ~3,300 frames/s?  GPU-side throughput of the drop-in's execution shape with the
# pruned search (the mirror's default), and the CLI with 15 / 8 / 4 entropy workers (a warm-up run first: the first process on a fresh box is slow)
mkdir -p gpurun_out
{
python scripts/slot_stream_probe.py 16 16 1 48 1
python scripts/slot_stream_probe.py 32 32 1 48 1
} 2>&1 | tee gpurun_out/r4b_slot_probe.txt
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
  s=$(date +%s.%N)
  env $1 LD_LIBRARY_PATH=video-encoder_b200 tools/b2_encode --size 1920x1080 --fps 60 --quality 26 --gop 32 $2 $IN /dev/shm/b2_probe_out.h264 > gpurun_out/cli_probe.out 2> gpurun_out/cli_probe.err || { echo "failed: $2"; tail -3 gpurun_out/cli_probe.err; }
  e=$(date +%s.%N)
  python -c "dt=$e-$s; print('%-26s b2_encode %-30s process %.2f s | loop: %s' % ('$1', '$2', dt, open('gpurun_out/cli_probe.out').read().strip().splitlines()[-1]))"
}
{
run "WARMUP=1" "--preset slow --slots 16"
run "A=1" "--preset slow --slots 16"
run "A=1" "--preset slow --slots 32"
run "B2ENC_ENTROPY_THREADS=8" "--preset slow --slots 32"
run "B2ENC_ENTROPY_THREADS=4" "--preset slow --slots 32"
run "B2ENC_ME_PRUNE=0" "--preset slow --slots 16"
} | tee gpurun_out/r4b_cli.txt
rm -f $IN /dev/shm/b2_probe_out.h264
