#!/bin/bash
# the bench-path parity tests must FAIL when an inter-stream wait is removed (engine test hook B2ENC_TEST_FAULT)
mkdir -p gpurun_out
{
echo "== no fault: expected to pass"
timeout 600 python -m pytest tests/test_bench_path.py -m gpu -q -x 2>&1 | tail -2
for f in no_h2d_wait no_d2h_wait; do
  echo "== B2ENC_TEST_FAULT=$f (no_h2d_wait is expected to FAIL: K0 reads ring entries before their upload has landed; no_d2h_wait only races when the copy-out lags two steps behind)"
  B2ENC_TEST_FAULT=$f timeout 600 python -m pytest tests/test_bench_path.py -m gpu -q -x 2>&1 | tail -4
done
} | tee gpurun_out/r2_fault_injection.log
