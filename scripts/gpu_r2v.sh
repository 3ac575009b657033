#!/bin/bash
mkdir -p gpurun_out
B2_FUZZ_PRUNE=1 timeout 600 python scripts/gpu_fuzz.py 150 21 > gpurun_out/r2v_fuzz_prune.txt 2>&1; tail -2 gpurun_out/r2v_fuzz_prune.txt
timeout 300 python scripts/gpu_fuzz.py 80 22 > gpurun_out/r2v_fuzz_mixed.txt 2>&1; tail -2 gpurun_out/r2v_fuzz_mixed.txt
export B2_ME_PRUNE=1
ncu --set full --clock-control none --import-source on -k regex:'sea_kernel|k1a_' -s 2 -c 2 -f -o /tmp/prof_r2v python scripts/ncu_target.py > gpurun_out/r2v_ncu.log 2>&1
ncu -i /tmp/prof_r2v.ncu-rep --page raw --csv > gpurun_out/r2v_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_r2v.ncu-rep --page source --csv -k regex:'sea_kernel' > gpurun_out/r2v_sea_source.csv 2>/dev/null
python scripts/ncu_phase_shares.py gpurun_out/r2v_sea_source.csv
