#!/bin/bash
# round 2, first GPU pass: whole GPU suite, the reference call sequence through the C CLI and the Python probe, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > gpurun_out/r2a_env.txt; nproc >> gpurun_out/r2a_env.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r2a_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 600 bash scripts/cli_probe.sh 1024 > gpurun_out/r2a_cli.log 2>&1; tail -6 gpurun_out/r2a_cli.log
timeout 600 python scripts/dropin_probe.py > gpurun_out/r2a_dropin.log 2>&1; tail -12 gpurun_out/r2a_dropin.log
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/r2a_bench_conn8.json 2>> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench_conn8.json
