#!/bin/bash
mkdir -p gpurun_out
{
python scripts/frame_latency_probe.py 1920 1080 32 1
B2_K8_WAVEFRONT=1 python scripts/frame_latency_probe.py 1920 1080 32 1
python scripts/frame_latency_probe.py 1280 720 16 1
python scripts/frame_latency_probe.py 3840 2160 32 1
} > gpurun_out/r2c_latency.log 2>&1
cat gpurun_out/r2c_latency.log
python -m pytest tests/test_engine_parity.py -m gpu -q -x -k "deblock" 2>&1 | tail -2
