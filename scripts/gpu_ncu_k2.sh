#!/bin/bash
mkdir -p gpurun_out
python scripts/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k2_me_subpel' -s 0 -c 1 -o gpurun_out/prof_k2 -f python scripts/ncu_target.py > gpurun_out/ncu_k2.log 2>&1
tail -n 2 gpurun_out/ncu_k2.log
