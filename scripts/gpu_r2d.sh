#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_full_size.py tests/test_bench_path.py tests/test_dropin.py tests/test_fuzz_parity.py -m gpu -q -x > gpurun_out/r2d_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_tests.log; tail -4 gpurun_out/r2d_tests.log
B2_K8_ROW_WARPS=16 timeout 600 python -m pytest tests/test_engine_parity.py tests/test_full_size.py -m gpu -q -x -k "deblock or c3 or c4" 2>&1 | tail -2
{
python scripts/frame_latency_probe.py 1920 1080 32 1
B2_K8_ROW_WARPS=16 python scripts/frame_latency_probe.py 1920 1080 32 1
python scripts/frame_latency_probe.py 3840 2160 32 1
for cfg in "16 16 1" "32 32 1" "16 16 0"; do
  timeout 300 python scripts/slot_stream_probe.py $cfg 48
  B2_K8_ROW_WARPS=16 timeout 300 python scripts/slot_stream_probe.py $cfg 48
done
} > gpurun_out/r2d_probe.log 2>&1
cat gpurun_out/r2d_probe.log
