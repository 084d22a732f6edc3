#!/bin/bash
mkdir -p gpurun_out
N=${N:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
cut -c1-260 gpurun_out/bench_n$N.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print(d['value'], d['ms_per_step'], d.get('c5'))"
