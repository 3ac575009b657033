#!/bin/bash
# 2-GPU box: closed-GOP sharding of ONE stream inside libb2enc.so (T5), host-copy ceiling, CLI with --devices
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2m2_env.txt; nproc >> gpurun_out/r2m2_env.txt
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_sharding.py -m gpu -q -x -s > gpurun_out/r2m2_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m2_tests.log
tail -6 gpurun_out/r2m2_tests.log
for n in 1 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 scripts/pcie_ceiling.py 2>/dev/null | tail -1
done | tee gpurun_out/r2m2_pcie.log
