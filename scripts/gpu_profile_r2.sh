#!/bin/bash
# round-2 profile refresh: launch lists (per-launch device time) and ncu --set full captures of every kernel, for the named
# path and for the all-features configuration; each ncu run only after the plain run of the same command exited 0.
# The reports are turned into CSV pages on the box (raw page of everything, per-instruction source page of K1) and deleted:
# gpurun_out/ may only carry 64 MiB back.
mkdir -p gpurun_out
for F in 0 1; do
  export B2_ALL_FEATURES=$F
  python scripts/ncu_target.py > gpurun_out/ncu_plain_r2_f$F.log 2>&1 || { echo "plain run failed (features=$F)"; tail -5 gpurun_out/ncu_plain_r2_f$F.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_f$F.csv python scripts/ncu_target.py > gpurun_out/ncu_launches_r2_f$F.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k0_|k1_|k2_|k3_|k5_|k6_|k7_|k8_|k9' -s 9 -c 13 -f -o /tmp/prof_r2_f$F python scripts/ncu_target.py > gpurun_out/ncu_full_r2_f$F.log 2>&1
  ncu -i /tmp/prof_r2_f$F.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_raw_f$F.csv 2>/dev/null
  if [ $F = 0 ]; then ncu -i /tmp/prof_r2_f0.ncu-rep --page source --csv -k regex:'k1_me_fullpel' > gpurun_out/r2_k1_source_page.csv 2>/dev/null; fi
  tail -n 2 gpurun_out/ncu_launches_r2_f$F.log; tail -n 2 gpurun_out/ncu_full_r2_f$F.log
done
unset B2_ALL_FEATURES
# the launch list of the bench command itself (share of K1 in the step must agree with bench.py's live number)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin --no-verify > gpurun_out/bench_for_launches.json 2>/dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 320 --csv --log-file gpurun_out/launches_r2_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin --no-verify > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out | grep -E "r2_ncu|r2_k1|launches_r2"; du -sh gpurun_out
