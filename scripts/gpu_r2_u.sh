#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_batched.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2u_pytest.txt
B="--no-extras --no-cpu-baseline --e2e-steps 10"
run() { timeout 200 python bench.py --config c3 --steps 20 --warmup 5 $B $1 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step')"; }
{
echo "== bare default"; run
echo "== wrapped default"; run --wrapped-main
echo "== wrapped no replay"; QS_WRAP_REPLAY=0 run --wrapped-main
echo "== wrapped lockstep"; run "--wrapped-main --lockstep"
} 2>&1 | tee gpurun_out/r2u_ab.txt
for tool in memcheck racecheck; do echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/gpu_sanitize.py 2>&1 | tail -6; done 2>&1 | tee gpurun_out/r2u_sanitize.txt
