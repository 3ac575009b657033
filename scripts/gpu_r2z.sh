#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_me_fullpel.py -m gpu -q -x -k "pruned or block_sums" 2>&1 | tail -2
timeout 900 python scripts/k1_pde_probe.py 2>&1 | tee gpurun_out/r2z_pde_probe.txt
B2_K1_PRUNE_ROWS=fine python scripts/k1_prune_probe.py child 2>&1 | cut -c1-330 | head -2
