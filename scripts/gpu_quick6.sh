#!/bin/bash
mkdir -p gpurun_out
for cfg in ${B2_CFGS:-"32 8" "64 8"}; do
  set -- ${cfg/_/ }
  B2_BENCH_SLOTS=$1 B2_BENCH_STREAMS=$2 python bench.py --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
  python - "$cfg" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print(repr(sys.argv[1]), d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
PY
done
