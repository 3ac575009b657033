#!/bin/bash
# whole-step throughput with K1 residency variants (scripts/k1_variants.sh): does leaving room on the SM for the other
# stream groups' kernels pay more than K1's own third CTA?
mkdir -p gpurun_out
for lib in build_variants/*.so; do
  B2ENC_LIB=$PWD/$lib python bench.py --no-cpu-baseline ${B2_VAR_ARGS:-} > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err
  python - "$lib" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/bench_var.json').read().strip().splitlines()[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'K1 frac', d['roofline']['frac'], 'K1 ms', d['kernel_ms_per_step_alone']['K1 full-pel SAD'])
PY
done 2>&1 | tee gpurun_out/k1_coreside.txt
