#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_full_size.py tests/test_bench_path.py tests/test_dropin.py tests/test_fuzz_parity.py tests/test_cli.py -m gpu -q -x > gpurun_out/r2e_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_tests.log; tail -4 gpurun_out/r2e_tests.log
{
python scripts/frame_latency_probe.py 1920 1080 32 1
python scripts/frame_latency_probe.py 1280 720 16 1
python scripts/frame_latency_probe.py 3840 2160 32 1
for cfg in "16 16 1" "32 32 1" "16 16 0" "8 8 1"; do timeout 300 python scripts/slot_stream_probe.py $cfg 48; done
} > gpurun_out/r2e_probe.log 2>&1
cat gpurun_out/r2e_probe.log
timeout 600 bash scripts/cli_probe.sh 2048 > gpurun_out/r2e_cli.log 2>&1; tail -5 gpurun_out/r2e_cli.log
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-verify --deblock 1 > gpurun_out/r2e_bench_deblock.json 2> gpurun_out/r2e_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2e_bench_deblock.json').read().strip().splitlines()[-1]); print('bench deblock=1', d['value'], d['e2e']['value'], d['kernel_ms_per_step_alone'], 'dropin', d['dropin']['value'], d['dropin']['first_output_after_pictures'], 'named', d['dropin_named_path']['value'])"
