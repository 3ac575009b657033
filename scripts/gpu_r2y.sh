#!/bin/bash
# end-of-round validation on one GPU with the final kernels (pruned search on by default behind the x264 mirror)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2y_tests.log; tail -3 gpurun_out/r2y_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python scripts/gpu_fuzz_dropin.py 60 31 > gpurun_out/r2y_fuzz_dropin.txt 2>&1; tail -1 gpurun_out/r2y_fuzz_dropin.txt
show() { python -c "
import json,sys; d=json.loads(open('$1').read().strip().splitlines()[-1]); p=d.get('pruned') or {}
print('$2', 'value', d['value'], 'e2e', d['e2e']['value'], 'verified', d['verified'], 'frac', d['roofline']['frac'], 'dropin', (d.get('dropin') or {}).get('value'), (d.get('dropin_named_path') or {}).get('value'), '| pruned value', p.get('value'), 'e2e', p.get('e2e'), 'verified', p.get('verified'), 'executed', p.get('k1_executed_fraction'), 'k1 ms', p.get('k1_ms_per_step_alone'), 'vs', p.get('k1_ms_per_step_alone_exhaustive'))"; }
timeout 900 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -2 gpurun_out/r2y_bench.err; show gpurun_out/r2y_bench.json c3
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-dropin --deblock 1 --transform8x8 1 --partitions 1 > gpurun_out/r2y_bench_allfeatures.json 2> gpurun_out/r2y.err; show gpurun_out/r2y_bench_allfeatures.json all-features
timeout 600 python bench.py --workload c4 --steps 16 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/r2y_bench_c4.json 2> gpurun_out/r2y.err; show gpurun_out/r2y_bench_c4.json c4
B2_ME_PRUNE=1 python scripts/ncu_target.py > /dev/null 2>&1 && B2_ME_PRUNE=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2y_launches_pruned.csv python scripts/ncu_target.py > /dev/null 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2y_launches_pruned.csv 2>/dev/null | head -30
