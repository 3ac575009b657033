#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r3e_bench_n2.json 2> gpurun_out/r3e.err; echo "exit $?"; tail -2 gpurun_out/r3e.err | cut -c1-200
python -c "
import json; d=json.loads(open('gpurun_out/r3e_bench_n2.json').read().strip().splitlines()[-1]); p=d['pruned']; print('n2', d['n_gpus'], d['value'], d['e2e']['value'], d['verified'], d['roofline']['frac'], '| pruned', p['value'], p['e2e'], p['verified'], p['k1_executed_fraction'])"
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/r3e_bench_n1.json 2> gpurun_out/r3e.err; echo "exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r3e_bench_n1.json').read().strip().splitlines()[-1]); p=d['pruned']; print('n1', d['n_gpus'], d['value'], d['e2e']['value'], d['verified'], d['roofline']['frac'], '| pruned', p['value'], p['e2e'], p['verified'], p['k1_executed_fraction'])"
