#!/bin/bash
# first GPU contact: parity of K1 + integer peak + K1 timing
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python scripts/k1_probe.py > gpurun_out/k1_probe.log 2>&1
cat gpurun_out/k1_probe.log
