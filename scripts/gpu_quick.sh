#!/bin/bash
# quick check: GPU tests + all-features bench (kernel times alone)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --deblock 1 --transform8x8 1 --partitions 1 --no-cpu-baseline > gpurun_out/bench_quick_all.json 2> gpurun_out/bench_quick_all.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick_all.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms_per_step_alone'])
PY
