#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_me_fullpel.py -m gpu -q -x -k "pruned or block_sums" 2>&1 | tail -5
timeout 900 python scripts/k1_prune_probe.py 2>&1 | tee gpurun_out/r2r_k1_prune_probe.txt
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_bench_path.py tests/test_dropin.py -m gpu -q -x -k "pruned or pruning or call_sequence" 2>&1 | tail -5
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; tail -3 gpurun_out/r2r_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2r_bench.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['e2e']['value'], d['verified'], 'roofline', d['roofline']['frac']); print(json.dumps({k:v for k,v in d['pruned'].items() if k!='what'}))"
