#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for W in c2 c5; do python bench.py --workload $W --no-cpu-baseline > gpurun_out/bench_r1g_$W.json 2> gpurun_out/bench_r1g_$W.err; python - $W <<'PY'
import json, sys
d = json.loads(open('gpurun_out/bench_r1g_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
PY
done
