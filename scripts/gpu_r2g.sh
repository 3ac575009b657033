#!/bin/bash
# single-GPU experiments: C5 stream-group count, persistent K1 at +-16, CLI slot counts, fuzz
mkdir -p gpurun_out
run() { # name, env..., -- args
  name=$1; shift
  env "$@" > /dev/null 2>&1
}
for s in 4 8; do
  B2_BENCH_STREAMS=$s timeout 300 python bench.py --workload c5 --steps 40 --warmup 5 --no-cpu-baseline --no-dropin --no-verify 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 streams $s: value %.0f e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
for p in 0 1; do
  B2_K1_PERSISTENT=$p timeout 300 python bench.py --workload c2 --steps 12 --warmup 3 --no-cpu-baseline --no-dropin --no-verify 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 persistentK1 $p: value %.0f e2e %.0f K1 frac %.3f K0 frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['hbm']['frac']), d['kernel_ms_per_step_alone'])"
done
python - <<'PY'
import sys
sys.path.insert(0, "oracle")
import b2oracle as o
fr = [b"".join(p.tobytes() for p in o.synth_frame(1920, 1080, t)) for t in range(32)]
with open("/dev/shm/b2_1080p.yuv", "wb") as f:
    for i in range(3072): f.write(fr[i % 32] if (i // 32) % 2 == 0 else fr[31 - i % 32])
PY
for extra in "--slots 16" "--slots 24" "--slots 32" "--slots 32 --profile baseline"; do
  out=$(LD_LIBRARY_PATH=video-encoder_b200 timeout 300 tools/b2_encode --size 1920x1080 --fps 60 --quality 26 --gop 32 --preset slow $extra /dev/shm/b2_1080p.yuv /dev/shm/b2_out_x.h264 2>&1 | tail -1)
  echo "b2_encode 1080p $extra: $out"
done
rm -f /dev/shm/b2_1080p.yuv /dev/shm/b2_out_x.h264
timeout 600 python scripts/gpu_fuzz.py 300 2>&1 | tail -2
timeout 600 python scripts/gpu_fuzz_dropin.py 200 2>&1 | tail -2
