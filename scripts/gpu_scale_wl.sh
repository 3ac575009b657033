#!/bin/bash
# BASELINE.json configs C4 (2160p closed-GOP sharding) and C5 (64 live 720p streams = 8 per GPU) on N GPUs
N=${1:-8}
mkdir -p gpurun_out
for W in c4 c5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $W --steps 32 --warmup 3 > gpurun_out/bench_${W}_n$N.json 2> gpurun_out/bench_${W}_n$N.err
  python - $W $N <<'PY'
import json, sys
d = json.loads(open('gpurun_out/bench_%s_n%s.json' % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
print(sys.argv[1], 'n', d['n_gpus'], d['metric'], d['value'], 'e2e', d['e2e']['value'], d['ms_per_step'])
PY
done
