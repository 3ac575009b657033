#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_fault_injection.sh
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2j_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_tests.log; tail -3 gpurun_out/r2j_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['e2e']['value'], d['verified'], d['roofline']['frac'], 'dropin', d['dropin']['value'], d['dropin']['first_output_after_pictures'], 'named', d['dropin_named_path']['value'], 'cpu', d['cpu_baseline']['value'])"
timeout 600 python bench.py --steps 12 --warmup 3 --deblock 1 --transform8x8 1 --partitions 1 --no-cpu-baseline --no-dropin > gpurun_out/r2j_bench_allfeatures.json 2>> gpurun_out/r2j_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2j_bench_allfeatures.json').read().strip().splitlines()[-1]); print('all features', d['value'], d['e2e']['value'], d['verified'], d['kernel_ms_per_step_alone'])"
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reference arm', d['value'], d['cpu_baseline']['cores'])"
