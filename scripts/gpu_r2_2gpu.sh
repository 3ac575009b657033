#!/bin/bash
# 2-GPU box: closed-GOP sharding of ONE stream inside libb2enc.so (T5), host-copy ceiling, K7/K8 parity after the latest changes
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2m2_env.txt; nproc >> gpurun_out/r2m2_env.txt
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_sharding.py -m gpu -q -x -s > gpurun_out/r2m2_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m2_tests.log
tail -6 gpurun_out/r2m2_tests.log
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_full_size.py tests/test_dropin.py -m gpu -q -x 2>&1 | tail -3
for n in 1 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 scripts/pcie_ceiling.py 2>/dev/null | tail -1
done | tee gpurun_out/r2m2_pcie.log
python scripts/frame_latency_probe.py 1920 1080 32 1 2>&1 | grep "1 slot"
python scripts/frame_latency_probe.py 3840 2160 32 1 2>&1 | grep "1 slot"
