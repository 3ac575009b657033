#!/bin/bash
# round 2, second GPU pass: row-pipelined K8 parity + drop-in shape throughput (per-slot stream groups) with both K8 schedules
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_full_size.py tests/test_dropin.py tests/test_bench_path.py tests/test_fuzz_parity.py -m gpu -q -x > gpurun_out/r2b_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_tests.log
tail -4 gpurun_out/r2b_tests.log
{
for cfg in "16 16 0" "16 16 1" "16 4 1" "16 8 1" "32 32 1" "32 8 1" "8 8 1"; do
  timeout 300 python scripts/slot_stream_probe.py $cfg 48
  B2_K8_WAVEFRONT=1 timeout 300 python scripts/slot_stream_probe.py $cfg 48
done
} > gpurun_out/r2b_slot_probe.log 2>&1
cat gpurun_out/r2b_slot_probe.log
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-dropin --no-verify --deblock 1 > gpurun_out/r2b_bench_deblock.json 2> gpurun_out/r2b_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_deblock.json').read().strip().splitlines()[-1]); print('bench deblock=1', d['value'], d['e2e']['value'], d['kernel_ms_per_step_alone'])"
B2_K8_WAVEFRONT=1 timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-dropin --no-verify --deblock 1 > gpurun_out/r2b_bench_deblock_wf.json 2>> gpurun_out/r2b_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_deblock_wf.json').read().strip().splitlines()[-1]); print('bench deblock=1 wavefront', d['value'], d['e2e']['value'], d['kernel_ms_per_step_alone'])"
timeout 600 bash scripts/cli_probe.sh 1024 > gpurun_out/r2b_cli.log 2>&1; tail -5 gpurun_out/r2b_cli.log
