#!/bin/bash
mkdir -p gpurun_out
python scripts/k1_r16_variant_probe.py 2>&1 | tee gpurun_out/r2i_k1_r16_variants.log
