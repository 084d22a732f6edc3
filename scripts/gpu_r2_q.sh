#!/bin/bash
mkdir -p gpurun_out
B="--no-extras --no-cpu-baseline --e2e-steps 10"
run() { timeout 200 python bench.py --config $1 --steps 20 --warmup 5 $B $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step frac',round(d['roofline']['frac'],4))"; }
{
for c in c3 c5; do echo "== $c default"; run $c; done
for b in 32 64 128; do echo "== c4 QS_BLOCK=$b"; QS_BLOCK=$b run c4; done
echo "== c4 QS_BLOCK=64 lockstep"; QS_BLOCK=64 run c4 --lockstep
for b in 64 128 256; do echo "== c3 nobalance QS_BLOCK=$b"; QS_BALANCE=0 QS_BLOCK=$b run c3; done
} 2>&1 | tee gpurun_out/r2q_ab.txt
python -m pytest tests/test_gpu_api.py -m gpu -q -x -k "one_output_array or back_to_back" 2>&1 | tail -5 | tee gpurun_out/r2q_pytest.txt
for tool in memcheck racecheck synccheck; do echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/gpu_sanitize.py 2>&1 | tail -8; done 2>&1 | tee gpurun_out/r2q_sanitize.txt
