#!/bin/bash
# round 2, run A: GPU tests, the driver's bench invocation, A/B of the new launch / write-out paths
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a_pytest.txt
cat gpurun_out/r2a_pytest.txt
B="--no-extras --no-cpu-baseline --e2e-steps 10"
run() { name=$1; shift; echo "== $name: $*" | tee -a gpurun_out/r2a_ab.txt; timeout 300 env "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print(round(d['ms_per_step']*1e3,3),'us/step frac',round(d['roofline']['frac'],4),'blocks',d['timing']['blocks'],'clk',d['clocks'])
except Exception as e: print('FAILED', l[-600:])
" | tee -a gpurun_out/r2a_ab.txt; }
run default_s20 python bench.py --steps 20 --warmup 5 $B
run lockstep_s20 python bench.py --steps 20 --warmup 5 --lockstep $B
run vecstores_s20 QS_OBS_BULK=0 python bench.py --steps 20 --warmup 5 $B
run vecstores_lock QS_OBS_BULK=0 python bench.py --steps 20 --warmup 5 --lockstep $B
run handover_s20 QS_PDL=3 python bench.py --steps 20 --warmup 5 $B
run handover_lock QS_PDL=3 python bench.py --steps 20 --warmup 5 --lockstep $B
run long_default python bench.py --steps 20000 --warmup 64 $B
run c2_s20 python bench.py --config c2 --steps 20 --warmup 5 $B
run c4_s20 python bench.py --config c4 --steps 20 --warmup 5 $B
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_full.json 2> gpurun_out/r2a_bench_full.err
tail -c 3000 gpurun_out/r2a_bench_full.json
