#!/bin/bash
# pruned-search legs beside the exhaustive ones: all features, the other named workloads, and the C CLI with the switch on / off
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.loads(open('$1').read().strip().splitlines()[-1]); p=d.get('pruned') or {}
print('$2', 'value', d['value'], 'e2e', d['e2e']['value'], 'verified', d['verified'], 'frac', d['roofline']['frac'], '| pruned value', p.get('value'), 'e2e', p.get('e2e'), 'verified', p.get('verified'), 'executed', p.get('k1_executed_fraction'), 'k1 ms', p.get('k1_ms_per_step_alone'), 'vs', p.get('k1_ms_per_step_alone_exhaustive'))"; }
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-dropin --deblock 1 --transform8x8 1 --partitions 1 > gpurun_out/r2u_bench_allfeatures.json 2> gpurun_out/r2u.err; show gpurun_out/r2u_bench_allfeatures.json all-features
for wl in c2 c4 c5; do
  timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/r2u_bench_$wl.json 2> gpurun_out/r2u.err; show gpurun_out/r2u_bench_$wl.json $wl
done
IN=/dev/shm/b2_probe_1080p.yuv
python - "$IN" 3072 <<'PY'
import sys, os
sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import b2oracle as o
fr = [b"".join(p.tobytes() for p in o.synth_frame(1920, 1080, t)) for t in range(32)]
with open(sys.argv[1], "wb") as f:
    for i in range(int(sys.argv[2])): f.write(fr[i % 32] if (i // 32) % 2 == 0 else fr[31 - i % 32])
PY
for prune in 0 1; do for slots in 16 32; do
  B2ENC_ME_PRUNE=$prune LD_LIBRARY_PATH=video-encoder_b200 tools/b2_encode --size 1920x1080 --fps 60 --quality 26 --gop 32 --preset slow --slots $slots $IN /dev/shm/b2_probe_out.h264 > gpurun_out/cli_probe.out 2> gpurun_out/cli_probe.err || tail -3 gpurun_out/cli_probe.err
  echo "b2_encode 1080p --preset slow --slots $slots B2ENC_ME_PRUNE=$prune : $(tail -1 gpurun_out/cli_probe.out)  sha=$(sha256sum /dev/shm/b2_probe_out.h264 | cut -c1-16)"
done; done | tee gpurun_out/r2u_cli_prune.txt
rm -f $IN /dev/shm/b2_probe_out.h264
nproc
