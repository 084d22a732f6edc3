#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
tail -c 1500 gpurun_out/r2e_bench_n2.json; tail -3 gpurun_out/r2e_bench_n2.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open(f'gpurun_out/r2e_bench_n{n}.json').read().strip().splitlines()[-1])
        print(n, 'value', d['value'], 'us/step', d['ms_per_step'] * 1e3, 'e2e', d['e2e']['value'], 'blocks', d['timing']['blocks'], d['timing'].get('metrics_gather'), d.get('c5'), d['clocks'])
    except Exception as e:
        print(n, 'FAILED', e)
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 2>&1 | tail -2 | cut -c1-900
