#!/bin/bash
# the driver's N>1 launch of bench.py with the final code (both arms)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r3d_bench_n2.json 2> gpurun_out/r3d.err; echo "exit $?"; tail -2 gpurun_out/r3d.err
python -c "
import json; d=json.loads(open('gpurun_out/r3d_bench_n2.json').read().strip().splitlines()[-1]); print('n2', d['n_gpus'], d['value'], d['e2e']['value'], d['verified'], d['roofline']['frac'], d.get('pruned'))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
