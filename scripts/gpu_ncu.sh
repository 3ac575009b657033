#!/bin/bash
mkdir -p gpurun_out
python scripts/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python scripts/ncu_target.py > gpurun_out/ncu_launches.log 2>&1
python scripts/ncu_target.py > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k1_me_fullpel|k2_me_subpel|k5_decide|k7_intra|k3_intra|k0_convert' -s 10 -c 8 -o gpurun_out/prof_r1 python scripts/ncu_target.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_launches.log gpurun_out/ncu_full.log
ls -la gpurun_out/
