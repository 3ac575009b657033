#!/bin/bash
bash scripts/gpu_quick.sh
python bench.py --no-cpu-baseline --no-dropin --partitions 2 > gpurun_out/bench_p2.json 2> gpurun_out/bench_p2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_p2.json').read().strip().splitlines()[-1])
print('partitions 2:', d['value'], d['e2e']['value'], d['kernel_ms_per_step_alone'])
PY
