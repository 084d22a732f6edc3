#!/bin/bash
mkdir -p gpurun_out
B="--no-extras --no-cpu-baseline --e2e-steps 10"
run() { timeout 200 python bench.py --config $1 --steps 20 --warmup 5 $B $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step frac',round(d['roofline']['frac'],4))"; }
{
for kb in 120 100 70 48; do for c in c3 c5; do echo "== $c courier KB=$kb"; QS_BALANCE_KB=$kb run $c; done; done
echo "== c3 courier KB=70 lockstep"; QS_BALANCE_KB=70 run c3 --lockstep
for c in c3 c5; do echo "== $c QS_COURIER=0"; QS_COURIER=0 run $c; done
for c in c2 c4; do echo "== $c default"; run $c; done
} 2>&1 | tee gpurun_out/r2o_ab.txt
python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2o_pytest.txt
export QS_LIB=$PWD/tune/libquadswarm_tl.so
for kb in 120 70; do QS_BALANCE_KB=$kb QS_PREGEN_HEAD=1 timeout 200 python scripts/gpu_timeline.py c3 stagger; done 2>&1 | tee gpurun_out/r2o_timeline.txt
