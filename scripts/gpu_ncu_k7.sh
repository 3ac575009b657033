#!/bin/bash
mkdir -p gpurun_out
python scripts/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k7_intra' -c 2 -o gpurun_out/prof_k7 python scripts/ncu_target.py > gpurun_out/ncu_k7.log 2>&1
tail -n 3 gpurun_out/ncu_k7.log
