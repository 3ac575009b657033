#!/bin/bash
# One 8-GPU box, everything that needs several GPUs (row X1, X2, item 5 of VERDICT r1):
#  1. T5: one 4K stream over 1/2/4/8 GPUs inside libb2enc.so, byte-identical (tests/test_multi_gpu.py)
#  2. host <-> device copy ceiling with the bench's byte counts at 1/2/4/8 ranks
#  3. the reference's call sequence in C (tools/b2_encode: mmap -> b2_sws_scale -> b2_encoder_encode -> fwrite) with --devices 1/2/4/8
#  4. bench.py --workload c2 / c4 / c5 (and c3) at 1/2/4/8 GPUs
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2m8_env.txt; nproc >> gpurun_out/r2m8_env.txt; nvidia-smi topo -m >> gpurun_out/r2m8_env.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_sharding.py -m gpu -q -x -s > gpurun_out/r2m8_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m8_tests.log
tail -5 gpurun_out/r2m8_tests.log
for n in 1 2 4 8; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 scripts/pcie_ceiling.py 2>/dev/null | tail -1
done | tee gpurun_out/r2m8_pcie.log
# C CLI: 1080p and 4K raw input on tmpfs, 1/2/4/8 GPUs
python - <<'PY'
import sys, os
sys.path.insert(0, "oracle")
import b2oracle as o
for name, w, h, n in (("1080p", 1920, 1080, 2048), ("2160p", 3840, 2160, 384)):
    fr = [b"".join(p.tobytes() for p in o.synth_frame(w, h, t)) for t in range(32)]
    with open("/dev/shm/b2_%s.yuv" % name, "wb") as f:
        for i in range(n): f.write(fr[i % 32] if (i // 32) % 2 == 0 else fr[31 - i % 32])
PY
for cfg in "1080p 1920x1080" "2160p 3840x2160"; do
  set -- $cfg
  for n in 1 2 4 8; do
    out=$(LD_LIBRARY_PATH=video-encoder_b200 timeout 300 tools/b2_encode --size $2 --fps 60 --quality 26 --gop 32 --preset slow --devices $n /dev/shm/b2_$1.yuv /dev/shm/b2_out_$n.h264 2>&1 | tail -1)
    echo "b2_encode $1 --devices $n: $out  sha=$(sha256sum /dev/shm/b2_out_$n.h264 | cut -c1-16)"
  done
done | tee gpurun_out/r2m8_cli.log
for extra in "--slots 32 --devices 1" "--slots 32 --devices 2" "--slots 8 --devices 8" "--profile baseline --devices 2" "--profile baseline --devices 8"; do
  out=$(LD_LIBRARY_PATH=video-encoder_b200 timeout 300 tools/b2_encode --size 1920x1080 --fps 60 --quality 26 --gop 32 --preset slow $extra /dev/shm/b2_1080p.yuv /dev/shm/b2_out_x.h264 2>&1 | tail -1)
  echo "b2_encode 1080p $extra: $out"
done | tee -a gpurun_out/r2m8_cli.log
rm -f /dev/shm/b2_*.yuv /dev/shm/b2_out_*.h264
for wl in c2 c4 c5; do
  for n in 1 2 4 8; do
    if [ $n = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 bench.py"; fi
    timeout 300 $cmd --gpus $n --steps 8 --warmup 3 --workload $wl --no-cpu-baseline --no-dropin --no-verify 2>gpurun_out/r2m8_bench_${wl}_n$n.err | tail -1 > gpurun_out/r2m8_bench_${wl}_n$n.json
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2m8_bench_${wl}_n$n.json").read())
    print("${wl} n=$n value %.0f e2e %.0f ms/step %.3f K1 frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"]))
except Exception as e:
    print("${wl} n=$n FAILED", e)
PY
  done
done | tee gpurun_out/r2m8_bench_table.log
