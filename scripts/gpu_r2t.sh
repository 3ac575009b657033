#!/bin/bash
mkdir -p gpurun_out
B2_K1_PRUNE_ROWS=fine python scripts/k1_prune_probe.py child 2>&1 | cut -c1-330 | tee gpurun_out/r2t_probe.txt
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; tail -3 gpurun_out/r2t_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['e2e']['value'], d['verified'], 'roofline', d['roofline']['frac'], 'dropin', d['dropin']['value'], d['dropin_named_path']['value']); print(json.dumps({k:v for k,v in d['pruned'].items() if k!='what'}))"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2t_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2t_tests.log; tail -4 gpurun_out/r2t_tests.log
