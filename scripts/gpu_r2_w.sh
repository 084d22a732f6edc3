#!/bin/bash
mkdir -p gpurun_out
B="--no-extras --no-cpu-baseline --e2e-steps 10"
run() { timeout 200 python bench.py --config $1 --steps 20 --warmup 5 $B $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step')"; }
{
echo "== c3 bare"; run c3
echo "== c5 bare"; run c5
echo "== c3 wrapped"; run c3 --wrapped-main
echo "== c3 bare lockstep"; run c3 --lockstep
} 2>&1 | tee gpurun_out/r2w_ab.txt
python -m pytest tests/test_gpu_api.py tests/test_gpu_batched.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2w_pytest.txt
for tool in synccheck racecheck memcheck; do echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 10 python scripts/gpu_sanitize.py 2>&1 | grep -v "Host Frame\|Saved host" | tail -8; done 2>&1 | tee gpurun_out/r2w_sanitize.txt
