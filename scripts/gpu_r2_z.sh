#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_api.py tests/test_gpu_batched.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2z_pytest.txt
B="--no-extras --no-cpu-baseline --e2e-steps 10"
run() { timeout 200 python bench.py --config $1 --steps 20 --warmup 5 $B $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step')"; }
{
echo "== c2 default"; run c2
echo "== c2 QS_SPLIT=1"; QS_SPLIT=1 run c2
echo "== c3 default"; run c3
} 2>&1 | tee gpurun_out/r2z_ab.txt
