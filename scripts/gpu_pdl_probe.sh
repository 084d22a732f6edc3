#!/bin/bash
# Dev probe: c3 step time for the launch-chaining modes (QS_PDL), two repetitions each.
export OMP_NUM_THREADS=1 PYTHONUNBUFFERED=1
mkdir -p gpurun_out
out=gpurun_out/pdl_probe.txt
: > $out
for rep in 1 2; do for m in 0 2 4 3; do
    us=$(QS_PDL=$m timeout 100 python bench.py --config c3 --no-extras --no-cpu-baseline --e2e-steps 10 --steps 30000 --warmup 512 2>/dev/null | tail -1 | python -c "import sys,json; print('%.3f' % (json.loads(sys.stdin.read())['ms_per_step']*1e3))")
    echo "QS_PDL=$m c3 : $us us/step" | tee -a $out
done; done
