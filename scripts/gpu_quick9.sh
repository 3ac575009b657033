#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/cli_probe.sh 1024
timeout 600 python scripts/dropin_probe.py 2>&1 | tee gpurun_out/dropin_probe.txt
