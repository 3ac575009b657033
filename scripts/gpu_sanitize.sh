#!/bin/bash
mkdir -p gpurun_out
python scripts/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -5 gpurun_out/sanitize_plain.log; exit 1; }
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
