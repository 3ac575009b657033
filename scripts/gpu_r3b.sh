#!/bin/bash
mkdir -p gpurun_out
for cfg in "64 8" "64 16" "64 4" "96 8" "128 16"; do set -- $cfg
  B2_BENCH_SLOTS=$1 B2_BENCH_STREAMS=$2 timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-dropin --no-verify 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['pruned']; print('slots $1 groups $2: exhaustive', d['value'], d['e2e']['value'], '| pruned', p['value'], p['e2e'])"
done | tee gpurun_out/r3b_groups.txt
