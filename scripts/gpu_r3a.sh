#!/bin/bash
mkdir -p gpurun_out
for nt in 256 320 384 192; do echo "B2_K1_THREADS=$nt: $(B2_K1_THREADS=$nt python scripts/k1_pde_probe.py child 2>&1 | tail -1)"; done | tee gpurun_out/r3a_threads.txt
