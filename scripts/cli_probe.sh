#!/bin/bash
# throughput of tools/b2_encode (the reference's call sequence in C: fread -> picture -> b2_encoder_encode -> fwrite) at 1080p,
# raw I420 input on tmpfs so that the file read is a memory copy
mkdir -p gpurun_out
IN=/dev/shm/b2_probe_1080p.yuv
N=${1:-1024}
python - "$IN" "$N" <<'PY'
import sys, os
ROOT = os.getcwd()
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b2oracle as o
fr = [b"".join(p.tobytes() for p in o.synth_frame(1920, 1080, t)) for t in range(32)]
with open(sys.argv[1], "wb") as f:
    for i in range(int(sys.argv[2])): f.write(fr[i % 32] if (i // 32) % 2 == 0 else fr[31 - i % 32])
PY
ls -la $IN
for args in "--preset slow --slots 8" "--preset slow --slots 16" "--preset slow --profile baseline --slots 16" "--preset slow --8x8dct --partitions 2 --slots 16"; do
  s=$(date +%s.%N)
  LD_LIBRARY_PATH=video-encoder_b200 tools/b2_encode --size 1920x1080 --fps 60 --quality 26 --gop 32 $args $IN /dev/shm/b2_probe_out.h264 > gpurun_out/cli_probe.out 2> gpurun_out/cli_probe.err || { echo "failed: $args"; tail -3 gpurun_out/cli_probe.err; }
  e=$(date +%s.%N)
  python -c "import os,sys; n=$N; dt=$e-$s; print('b2_encode %-52s process %.2f s (start-up included) | loop: %s' % ('$args', dt, open('gpurun_out/cli_probe.out').read().strip().splitlines()[-1]))"
done | tee gpurun_out/cli_probe.txt
rm -f $IN /dev/shm/b2_probe_out.h264
