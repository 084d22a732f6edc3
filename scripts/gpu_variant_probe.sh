#!/bin/bash
# time the c3 workload with several tuning builds (QS_LIB); mode 2 = grid-wide wait, 3 = per-block hand-over
export OMP_NUM_THREADS=1 PYTHONUNBUFFERED=1
mkdir -p gpurun_out; : > gpurun_out/variant_probe.txt
for rep in 1 2; do
for lib in "" tune/lib_v1.so tune/lib_v2.so tune/lib_v3.so; do
  for m in 2 3; do
    if [ "$m" = 3 ] && [ -n "$lib" ] && [ "$lib" != tune/lib_v3.so ]; then continue; fi
    r=$(QS_PDL=$m QS_LIB=${lib:+$PWD/$lib} timeout 120 python bench.py --steps 30000 --warmup 1024 --no-cpu-baseline --no-extras --e2e-steps 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3), 'us', round(d['roofline']['frac'],4))")
    echo "lib=${lib:-default} QS_PDL=$m : $r" | tee -a gpurun_out/variant_probe.txt
  done
done
done
