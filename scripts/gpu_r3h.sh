#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 4 --steps 12 --warmup 3 > gpurun_out/r3h_bench_n4.json 2> gpurun_out/r3h.err; echo "exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r3h_bench_n4.json').read().strip().splitlines()[-1]); p=d['pruned']; print('n4', d['n_gpus'], d['value'], d['e2e']['value'], d['verified'], d['roofline']['frac'], '| pruned', p['value'], p['e2e'], p['verified'], p['k1_executed_fraction'])"
nproc
