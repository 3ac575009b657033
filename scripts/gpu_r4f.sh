#!/bin/bash
# round 2, session 3: 4K through the C CLI with the reworked host stage (B2ENC_STATS), 1,024 frames, default 32 GOP slots and 16
mkdir -p gpurun_out
python - <<'PY'
import sys
sys.path.insert(0, "oracle")
import b2oracle as o
fr = [b"".join(p.tobytes() for p in o.synth_frame(3840, 2160, t)) for t in range(32)]
with open("/dev/shm/b2_2160p.yuv", "wb") as f:
    for i in range(1024): f.write(fr[i % 32] if (i // 32) % 2 == 0 else fr[31 - i % 32])
PY
run() {
  env $1 B2ENC_STATS=1 LD_LIBRARY_PATH=video-encoder_b200 timeout 100 tools/b2_encode --size 3840x2160 --fps 60 --quality 26 --gop 32 --preset slow $2 /dev/shm/b2_2160p.yuv /dev/shm/b2_out_x.h264 > gpurun_out/cli_probe.out 2> gpurun_out/cli_probe.err
  echo "$1 b2_encode 2160p $2 | loop: $(tail -1 gpurun_out/cli_probe.out) | sha $(sha256sum /dev/shm/b2_out_x.h264 | cut -c1-12)"; grep "b2enc stats" gpurun_out/cli_probe.err
}
{
run "WARMUP=1" "--slots 16" | head -1
run "A=1" ""
run "A=1" "--slots 16"
run "A=1" "--profile baseline"
} | tee gpurun_out/r4f_cli_4k.txt
rm -f /dev/shm/b2_2160p.yuv /dev/shm/b2_out_x.h264
