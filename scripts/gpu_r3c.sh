#!/bin/bash
# 2-GPU box, final code (search pruning on by default behind the x264 mirror): T5 and the multi-device tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r3c_env.txt; nproc >> gpurun_out/r3c_env.txt
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_sharding.py -m gpu -q -x -s > gpurun_out/r3c_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3c_tests.log
tail -8 gpurun_out/r3c_tests.log
