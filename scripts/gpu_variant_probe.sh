#!/bin/bash
# time the c3 workload with several tuning builds (QS_LIB) and block sizes
export OMP_NUM_THREADS=1 PYTHONUNBUFFERED=1
for lib in "" tune/lib_v1.so tune/lib_v2.so tune/lib_v3.so tune/lib_v4.so; do
  for b in 32 64; do
    echo "== lib=${lib:-default} block=$b"
    QS_LIB=${lib:+$PWD/$lib} QS_BLOCK=$b timeout 120 python bench.py --steps 30000 --warmup 1024 --no-cpu-baseline --e2e-steps 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,3), 'us', round(d['roofline']['frac'],4))"
  done
done
