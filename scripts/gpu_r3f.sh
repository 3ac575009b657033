#!/bin/bash
# final validation + evidence with the end-of-round code
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3f_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3f_tests.log; tail -3 gpurun_out/r3f_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1
s=$(date +%s); timeout 900 python bench.py > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; e=$(date +%s); echo "default bench.py run: $((e-s)) s"; tail -2 gpurun_out/r3f_bench.err | cut -c1-200
python -c "
import json; d=json.loads(open('gpurun_out/r3f_bench.json').read().strip().splitlines()[-1]); p=d['pruned']; print('c3', d['value'], d['e2e']['value'], d['verified'], d['roofline']['frac'], 'dropin', d['dropin']['value'], d['dropin_named_path']['value'], d['dropin']['first_output_after_ms'], 'cpu', d['cpu_baseline']['value'], '| pruned', p['value'], p['e2e'], p['verified'], p['k1_executed_fraction'], p['k1_ms_per_step_alone'], 'launches', d['gpu_launches'])"
export B2_ME_PRUNE=1
ncu --set full --clock-control none --import-source on -k regex:'sea_kernel|k1a_' -s 2 -c 2 -f -o /tmp/prof_r3f python scripts/ncu_target.py > gpurun_out/r3f_ncu.log 2>&1
ncu -i /tmp/prof_r3f.ncu-rep --page raw --csv > gpurun_out/r3f_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_r3f.ncu-rep --page source --csv -k regex:'sea_kernel' > gpurun_out/r3f_sea_source.csv 2>/dev/null
python scripts/ncu_phase_shares.py gpurun_out/r3f_sea_source.csv
python scripts/ncu_summary.py gpurun_out/r3f_ncu_raw.csv 2>/dev/null | head -12
