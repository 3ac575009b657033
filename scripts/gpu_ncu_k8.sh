#!/bin/bash
mkdir -p gpurun_out
python scripts/ncu_target_k8.py > gpurun_out/ncu_k8_plain.log 2>&1 || { tail -5 gpurun_out/ncu_k8_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k8_deblock_rows' -s 1 -c 1 -o gpurun_out/prof_r2_k8 -f python scripts/ncu_target_k8.py > gpurun_out/ncu_k8.log 2>&1
tail -3 gpurun_out/ncu_k8.log; ls -la gpurun_out/prof_r2_k8*
