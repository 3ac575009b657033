#!/bin/bash
# the reference's call sequence in C (tools/b2_encode: mmap -> b2_sws_scale -> b2_encoder_encode -> fwrite), ONE stream, 1..N GPUs
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
python - <<'PY'
import sys
sys.path.insert(0, "oracle")
import b2oracle as o
for name, w, h, n in (("1080p", 1920, 1080, 4096), ("2160p", 3840, 2160, 512)):
    fr = [b"".join(p.tobytes() for p in o.synth_frame(w, h, t)) for t in range(32)]
    with open("/dev/shm/b2_%s.yuv" % name, "wb") as f:
        for i in range(n): f.write(fr[i % 32] if (i // 32) % 2 == 0 else fr[31 - i % 32])
PY
{
echo "host cores: $(nproc), GPUs: $NG"
for cfg in "1080p 1920x1080" "2160p 3840x2160"; do
  set -- $cfg
  for n in 1 2 4 8; do
    [ $n -le $NG ] || continue
    for extra in "" "--profile baseline"; do
      out=$(LD_LIBRARY_PATH=video-encoder_b200 timeout 300 tools/b2_encode --size $2 --fps 60 --quality 26 --gop 32 --preset slow --devices $n $extra /dev/shm/b2_$1.yuv /dev/shm/b2_out.h264 2>&1 | tail -1)
      echo "b2_encode $1 --devices $n $extra: $out  sha=$(sha256sum /dev/shm/b2_out.h264 | cut -c1-16)"
    done
  done
done
} | tee gpurun_out/r2_cli_scaling.log
rm -f /dev/shm/b2_*.yuv /dev/shm/b2_out.h264
