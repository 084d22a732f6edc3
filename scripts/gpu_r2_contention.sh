#!/bin/bash
# 6 probe processes + 2 heavy background processes share the GPU; three launch shapes of the 37-env chain, one after the other.
export OMP_NUM_THREADS=1 PYTHONUNBUFFERED=1
mkdir -p gpurun_out
: > gpurun_out/contention.txt
run() {   # name, seconds, env assignments...
  local name=$1 secs=$2; shift 2
  for k in 1 2; do env "$@" timeout 60 python scripts/gpu_contention_probe.py bg$k $secs 37 bg >> gpurun_out/contention.txt 2>&1 & done
  for k in 1 2 3 4 5 6; do env "$@" timeout 60 python scripts/gpu_contention_probe.py $name.$k $secs >> gpurun_out/contention.txt 2>&1 & done
  wait
}
echo "== split + hand-over flags (QS_PDL=3), the failing test's shape" >> gpurun_out/contention.txt
run split_ho 14 QS_PDL=3
echo "== single shape + hand-over flags (QS_PDL=3 QS_SPLIT=0)" >> gpurun_out/contention.txt
run single_ho 10 QS_PDL=3 QS_SPLIT=0
echo "== split + grid-wide wait (QS_PDL=2)" >> gpurun_out/contention.txt
run split_wait 10 QS_PDL=2
echo "== alone, split + hand-over" >> gpurun_out/contention.txt
QS_PDL=3 timeout 40 python scripts/gpu_contention_probe.py alone 5 >> gpurun_out/contention.txt 2>&1
cat gpurun_out/contention.txt
