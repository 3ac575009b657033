#!/bin/bash
# round 2, session 3: the drop-in with the reworked staging copy / result-copy pool / 32 default GOP slots: CLI with B2ENC_STATS, the drop-in GPU
# tests, and bench.py's drop-in legs (engine legs unchanged since r3g: short run, no verification / pruned leg / CPU baseline)
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
run "A=1" "--preset slow --slots 16"
run "B2ENC_SWS_THREADS=1" "--preset slow"
run "A=1" "--preset slow --profile baseline"
run "A=1" "--preset slow --8x8dct --partitions 2"
} | tee gpurun_out/r4d_cli_stats.txt
rm -f $IN /dev/shm/b2_probe_out.h264
timeout 75 python -m pytest tests/test_dropin.py -m gpu -x -q > gpurun_out/r4d_tests.log 2>&1; echo "pytest test_dropin exit $?"; tail -2 gpurun_out/r4d_tests.log
timeout 90 python bench.py --steps 10 --warmup 3 --no-verify --no-pruned-leg --no-cpu-baseline > gpurun_out/r4d_bench.json 2> gpurun_out/r4d_bench.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r4d_bench.json').read().strip().splitlines()[-1]); a=d['dropin']; b=d['dropin_named_path']; print('c3', d['value'], d['e2e']['value'], 'dropin', a['value'], a['gop_slots_per_gpu'], a['first_output_after_ms'], '|', b['value'], b['first_output_after_ms'])"
