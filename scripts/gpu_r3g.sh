#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
s=$(date +%s); timeout 900 python bench.py > gpurun_out/r3g_bench_$i.json 2> gpurun_out/r3g_bench.err; e=$(date +%s); echo "default bench.py run: $((e-s)) s"
python -c "
import json; d=json.loads(open('gpurun_out/r3g_bench_$i.json').read().strip().splitlines()[-1]); p=d['pruned']; a=d['dropin']; b=d['dropin_named_path']; print('c3', d['value'], d['e2e']['value'], d['verified'], d['roofline']['frac'], 'dropin', a['value'], a['first_output_after_ms'], '|', b['value'], b['first_output_after_ms'], 'cpu', d['cpu_baseline']['value'], '| pruned', p['value'], p['e2e'], p['verified'], p['k1_executed_fraction'], p['k1_ms_per_step_alone'])"
done
