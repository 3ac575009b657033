#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_me_fullpel.py tests/test_engine_parity.py tests/test_bench_path.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python scripts/k1_exppad_probe.py 2>&1 | tee gpurun_out/r2m_k1_exppad.txt
for wl in c3 c2; do
timeout 300 python bench.py --workload $wl --steps 12 --warmup 3 --no-cpu-baseline --no-dropin --no-verify 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl', d['value'], d['roofline'])"
done
