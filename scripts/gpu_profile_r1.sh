#!/bin/bash
# round-1 profile refresh: launch lists (per-launch device time) and ncu --set full captures of every kernel, for the named
# path and for the all-features configuration; each ncu run only after the plain run of the same command exited 0
mkdir -p gpurun_out
for F in 0 1; do
  export B2_ALL_FEATURES=$F
  python scripts/ncu_target.py > gpurun_out/ncu_plain_f$F.log 2>&1 || { echo "plain run failed (features=$F)"; tail -5 gpurun_out/ncu_plain_f$F.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_f$F.csv python scripts/ncu_target.py > gpurun_out/ncu_launches_f$F.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k0_|k1_|k2_|k3_|k5_|k6_|k7_|k8_|k9' -s 12 -c 12 -o gpurun_out/prof_r1_f$F python scripts/ncu_target.py > gpurun_out/ncu_full_f$F.log 2>&1
  tail -n 2 gpurun_out/ncu_launches_f$F.log; tail -n 2 gpurun_out/ncu_full_f$F.log
done
unset B2_ALL_FEATURES
python bench.py > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err; tail -c 300 gpurun_out/bench_r1z.json
python bench.py --deblock 1 --transform8x8 1 --partitions 1 --no-cpu-baseline > gpurun_out/bench_r1z_allfeatures.json 2> gpurun_out/bench_r1z_allfeatures.err
for W in c2 c4 c5; do python bench.py --workload $W --no-cpu-baseline > gpurun_out/bench_r1z_$W.json 2> gpurun_out/bench_r1z_$W.err; done
ls -la gpurun_out | tail -20
