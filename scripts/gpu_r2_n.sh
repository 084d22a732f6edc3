#!/bin/bash
mkdir -p gpurun_out
B="--no-extras --no-cpu-baseline --e2e-steps 10"
run() { timeout 200 python bench.py --config $1 --steps 20 --warmup 5 $B $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step frac',round(d['roofline']['frac'],4))"; }
{
for kb in 120 100 70; do for c in c3 c5 c4; do echo "== $c EARLY KB=$kb"; QS_EARLY_RELEASE=1 QS_BALANCE_KB=$kb run $c; done; done
echo "== c3 EARLY KB=100 lockstep"; QS_EARLY_RELEASE=1 QS_BALANCE_KB=100 run c3 --lockstep
echo "== c3 EARLY KB=100 NOBALANCE"; QS_EARLY_RELEASE=1 QS_BALANCE=0 run c3
} 2>&1 | tee gpurun_out/r2n_ab.txt
QS_EARLY_RELEASE=1 QS_BALANCE_KB=100 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2n_pytest.txt
