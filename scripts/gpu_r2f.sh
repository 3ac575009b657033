#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_full_size.py tests/test_bench_path.py tests/test_fuzz_parity.py -m gpu -q -x > gpurun_out/r2f_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_tests.log; tail -3 gpurun_out/r2f_tests.log
{
python scripts/frame_latency_probe.py 1920 1080 32 1
python scripts/frame_latency_probe.py 3840 2160 32 1
for cfg in "16 16 1" "32 32 1"; do timeout 300 python scripts/slot_stream_probe.py $cfg 48; done
} > gpurun_out/r2f_probe.log 2>&1
grep -E "1 slot|slots" gpurun_out/r2f_probe.log
