#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import sys, os
sys.path.insert(0, "oracle")
import b2oracle as o
fr = [b"".join(p.tobytes() for p in o.synth_frame(1920, 1080, t)) for t in range(16)]
with open("/dev/shm/p.yuv", "wb") as f:
    for i in range(128): f.write(fr[i % 16])
PY
for i in 1 2; do
s=$(date +%s.%N)
LD_LIBRARY_PATH=video-encoder_b200 tools/b2_encode --size 1920x1080 --fps 60 --quality 26 --gop 32 --preset slow --slots 8 /dev/shm/p.yuv /dev/shm/o.h264
e=$(date +%s.%N)
python -c "print('process %.2f s' % ($e-$s))"
done
s=$(date +%s.%N); python -c "
import ctypes
L=ctypes.CDLL('video-encoder_b200/libb2enc.so')
"; e=$(date +%s.%N); python -c "print('dlopen only %.2f s' % ($e-$s))"
s=$(date +%s.%N); python -c "
import ctypes
L=ctypes.CDLL('libcudart.so.12')
L.cudaFree(0)
"; e=$(date +%s.%N); python -c "print('cudaFree(0) %.2f s' % ($e-$s))"
