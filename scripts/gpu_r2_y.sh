#!/bin/bash
mkdir -p gpurun_out
B="--no-extras --no-cpu-baseline --e2e-steps 10"
run() { timeout 200 python bench.py --config $1 --steps 20 --warmup 5 $B $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3),'us/step')"; }
{
for m in 40 0 10 100 296 256 266; do for c in c3 c5; do echo "== $c QS_POLL=$m"; QS_POLL=$m run $c; done; done
} 2>&1 | tee gpurun_out/r2y_ab.txt
