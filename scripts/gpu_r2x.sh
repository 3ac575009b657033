#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_me_fullpel.py -m gpu -q -x -k "pruned" 2>&1 | tail -2
B2_K1_PRUNE_ROWS=fine python scripts/k1_prune_probe.py child 2>&1 | cut -c1-330 | tee gpurun_out/r2x_probe.txt
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; tail -3 gpurun_out/r2x_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['e2e']['value'], d['verified'], 'roofline', d['roofline']['frac']); print(json.dumps({k:v for k,v in d['pruned'].items() if k!='what'}))"
