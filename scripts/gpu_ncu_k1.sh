#!/bin/bash
mkdir -p gpurun_out
python scripts/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k1_me_fullpel|k2_me_subpel|k3_intra' -s 3 -c 3 -o gpurun_out/prof_k1 python scripts/ncu_target.py > gpurun_out/ncu_k1.log 2>&1
tail -n 3 gpurun_out/ncu_k1.log
