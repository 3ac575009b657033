#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "" "--deblock 1" "--deblock 1 --transform8x8 1 --partitions 1"; do
  python bench.py $cfg --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
  python - "$cfg" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print(repr(sys.argv[1]), d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], {k: v for k, v in d['kernel_ms_per_step_alone'].items() if k[:2] in ('K7', 'K8', 'K3', 'K1')})
PY
done
