#!/bin/bash
mkdir -p gpurun_out
export QS_LIB=$PWD/tune/libquadswarm_tl.so
QS_PDL=2 QS_BALANCE=1 timeout 200 python scripts/gpu_timeline.py c3 > gpurun_out/r2c_tl_c3_pdl2_lock_bal.txt 2>&1; head -1 gpurun_out/r2c_tl_c3_pdl2_lock_bal.txt; tail -9 gpurun_out/r2c_tl_c3_pdl2_lock_bal.txt
QS_PDL=2 QS_PREGEN=16 timeout 200 python scripts/gpu_timeline.py c3 stagger > gpurun_out/r2c_tl_c3_pdl2_stagger.txt 2>&1; head -1 gpurun_out/r2c_tl_c3_pdl2_stagger.txt; tail -30 gpurun_out/r2c_tl_c3_pdl2_stagger.txt | grep "reset in\|median phase"
unset QS_LIB
B="--no-extras --no-cpu-baseline --e2e-steps 10"
for a in "" "--lockstep"; do for pdl in 2 3; do for bal in 0 1; do echo "== pdl $pdl bal $bal $a"; QS_BALANCE=$bal QS_PDL=$pdl timeout 200 python bench.py --steps 20 --warmup 5 $a $B 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step frac',round(d['roofline']['frac'],4))"; done; done; done 2>&1 | tee gpurun_out/r2c_ab.txt
for c in c2 c4 c5; do for bal in 0 1; do echo "== $c bal $bal"; QS_BALANCE=$bal timeout 200 python bench.py --config $c --steps 20 --warmup 5 $B 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step frac',round(d['roofline']['frac'],4))"; done; done 2>&1 | tee -a gpurun_out/r2c_ab.txt
