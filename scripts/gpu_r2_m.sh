#!/bin/bash
mkdir -p gpurun_out
B="--no-extras --no-cpu-baseline --e2e-steps 10"
run() { timeout 200 python bench.py --config $1 --steps 20 --warmup 5 $B $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step frac',round(d['roofline']['frac'],4))"; }
{
for c in c3 c2 c4 c5; do echo "== $c default"; run $c; done
echo "== c3 lockstep"; run c3 --lockstep
for kb in 100 70; do for c in c3 c5; do echo "== $c QS_BALANCE_KB=$kb"; QS_BALANCE_KB=$kb run $c; done; done
echo "== c3 lockstep QS_BALANCE_KB=100"; QS_BALANCE_KB=100 run c3 --lockstep
echo "== c3 QS_PDL=2 KB=100"; QS_PDL=2 QS_BALANCE_KB=100 run c3
echo "== c3 QS_PDL=2"; QS_PDL=2 run c3
echo "== c3 QS_PREGEN_HEAD=1"; QS_PREGEN_HEAD=1 run c3
} 2>&1 | tee gpurun_out/r2m_ab.txt
