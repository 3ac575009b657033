#!/bin/bash
mkdir -p gpurun_out
for wf in 0 1 auto; do
  if [ $wf = auto ]; then unset B2_K8_WAVEFRONT; else export B2_K8_WAVEFRONT=$wf; fi
  timeout 600 python bench.py --steps 12 --warmup 3 --deblock 1 --transform8x8 1 --partitions 1 --no-cpu-baseline --no-dropin --no-verify 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('all features, K8 wavefront=$wf:', d['value'], d['e2e']['value'], 'K8 alone', d['kernel_ms_per_step_alone']['K8 deblock'])"
  timeout 600 python bench.py --steps 12 --warmup 3 --deblock 1 --no-cpu-baseline --no-dropin --no-verify 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('deblock only, K8 wavefront=$wf:', d['value'], d['e2e']['value'])"
done 2>&1 | tee gpurun_out/r2k_k8_schedule_in_bench.log
unset B2_K8_WAVEFRONT
timeout 300 python -m pytest tests/test_engine_parity.py tests/test_bench_path.py -m gpu -q -x 2>&1 | tail -2
