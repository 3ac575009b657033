#!/bin/bash
mkdir -p gpurun_out
for lib in build_variants/libb2enc_prune_*.so; do
  echo "== $lib"
  B2ENC_LIB=$PWD/$lib B2_K1_PRUNE_ROWS=fine timeout 300 python scripts/k1_prune_probe.py child 2>&1 | grep "1920x1088" | cut -c1-330
done | tee gpurun_out/r2q_prune_strip_variants.txt
export B2_ME_PRUNE=1
python scripts/ncu_target.py > gpurun_out/r2q_plain.log 2>&1 || { tail -5 gpurun_out/r2q_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'sea_kernel|k1a_' -c 4 -f -o /tmp/prof_r2q python scripts/ncu_target.py > gpurun_out/r2q_ncu.log 2>&1
tail -2 gpurun_out/r2q_ncu.log
ncu -i /tmp/prof_r2q.ncu-rep --page raw --csv > gpurun_out/r2q_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_r2q.ncu-rep --page source --csv -k regex:'sea_kernel' > gpurun_out/r2q_sea_source.csv 2>/dev/null
ls -la gpurun_out | grep r2q
