#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin.py tests/test_cli.py tests/test_abi.py -m gpu -q -x 2>&1 | tail -3
timeout 600 bash scripts/cli_probe.sh 3072 > gpurun_out/r2h_cli.log 2>&1; tail -5 gpurun_out/r2h_cli.log
python - <<'PY'
import sys
sys.path.insert(0, "oracle")
import b2oracle as o
fr = [b"".join(p.tobytes() for p in o.synth_frame(3840, 2160, t)) for t in range(32)]
with open("/dev/shm/b2_2160p.yuv", "wb") as f:
    for i in range(512): f.write(fr[i % 32] if (i // 32) % 2 == 0 else fr[31 - i % 32])
PY
for extra in "" "--profile baseline"; do
  out=$(LD_LIBRARY_PATH=video-encoder_b200 timeout 300 tools/b2_encode --size 3840x2160 --fps 60 --quality 26 --gop 32 --preset slow $extra /dev/shm/b2_2160p.yuv /dev/shm/b2_out_x.h264 2>&1 | tail -1)
  echo "b2_encode 2160p $extra: $out"
done
rm -f /dev/shm/b2_2160p.yuv /dev/shm/b2_out_x.h264
timeout 300 python scripts/dropin_probe.py 2>&1 | grep "slots 16"
