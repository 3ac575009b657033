#!/bin/bash
mkdir -p gpurun_out
for cfg in "--partitions 2" "--deblock 1 --transform8x8 1 --partitions 2"; do
  python bench.py $cfg --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
  python - "$cfg" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print(repr(sys.argv[1]), d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'], d['kernel_ms_per_step_alone'])
open('gpurun_out/bench_r1f_' + sys.argv[1].replace('--', '').replace(' ', '_') + '.json', 'w').write(json.dumps(d))
PY
done
