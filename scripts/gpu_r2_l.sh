#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ticked_obstacle" 2>&1 | tail -15 | tee gpurun_out/r2l_new.txt
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2l_pytest.txt
B="--no-extras --no-cpu-baseline --e2e-steps 10"
for c in c3 c2 c4 c5; do echo "== $c"; timeout 200 python bench.py --config $c --steps 20 --warmup 5 $B 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step frac',round(d['roofline']['frac'],4))"; done 2>&1 | tee gpurun_out/r2l_ab.txt
echo "== c3 lockstep"; timeout 200 python bench.py --config c3 --steps 20 --warmup 5 $B --lockstep 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step frac',round(d['roofline']['frac'],4))" | tee -a gpurun_out/r2l_ab.txt
for pdl in 2 3; do echo "== c3 QS_PDL=$pdl"; QS_PDL=$pdl timeout 200 python bench.py --config c3 --steps 20 --warmup 5 $B 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step frac',round(d['roofline']['frac'],4))"; done 2>&1 | tee -a gpurun_out/r2l_ab.txt
