#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2l_tests.log; tail -3 gpurun_out/r2l_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python scripts/gpu_fuzz.py 120 7 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['e2e']['value'], d['e2e']['d2h_bytes_per_step'], d['verified'], 'dropin', d['dropin']['value'], d['dropin']['first_output_after_ms'], 'named', d['dropin_named_path']['value'])"
bash scripts/gpu_profile_r2.sh
