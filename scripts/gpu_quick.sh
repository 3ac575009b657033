#!/bin/bash
# GPU parity tests + one default bench line without the CPU baseline (development loop; scripts/gpu_verify.sh is the round checkpoint)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'K0', d['roofline']['hbm'], d['kernel_ms_per_step_alone'])
PY
