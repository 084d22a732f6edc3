#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2i_pytest.txt
B="--no-extras --no-cpu-baseline --e2e-steps 10"
for a in "" "--lockstep"; do for pdl in 2 3; do echo "== c3 pdl $pdl $a"; QS_PDL=$pdl timeout 200 python bench.py --steps 20 --warmup 5 $a $B 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step frac',round(d['roofline']['frac'],4))"; done; done 2>&1 | tee gpurun_out/r2i_ab.txt
CFG=c3 bash scripts/gpu_profile.sh > gpurun_out/r2i_profile.log 2>&1
tail -5 gpurun_out/r2i_profile.log
