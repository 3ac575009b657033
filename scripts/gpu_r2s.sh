#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_me_fullpel.py -m gpu -q -x -k "pruned or block_sums" 2>&1 | tail -3
B2_K1_PRUNE_ROWS=fine python scripts/k1_prune_probe.py child 2>&1 | cut -c1-330 | tee gpurun_out/r2s_probe.txt
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; tail -3 gpurun_out/r2s_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['e2e']['value'], d['verified'], 'roofline', d['roofline']['frac']); print(json.dumps({k:v for k,v in d['pruned'].items() if k!='what'}))"
export B2_ME_PRUNE=1
ncu --set full --clock-control none --import-source on -k regex:'sea_kernel' -s 1 -c 1 -f -o /tmp/prof_r2s python scripts/ncu_target.py > gpurun_out/r2s_ncu.log 2>&1
ncu -i /tmp/prof_r2s.ncu-rep --page raw --csv > gpurun_out/r2s_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_r2s.ncu-rep --page source --csv > gpurun_out/r2s_sea_source.csv 2>/dev/null
python scripts/ncu_phase_shares.py gpurun_out/r2s_sea_source.csv
