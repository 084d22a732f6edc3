#!/bin/bash
# Last GPU call of round 2: the contention probe again with the side stream ordered after the default stream (the fix) and
# without (control), the three tests that were changed, and the bench command after its one-line change.
export OMP_NUM_THREADS=1 PYTHONUNBUFFERED=1
mkdir -p gpurun_out
: > gpurun_out/contention2.txt
run() {   # name, seconds, extra probe arg, env assignments...
  local name=$1 secs=$2 extra=$3; shift 3
  for k in 1 2; do env "$@" timeout 60 python scripts/gpu_contention_probe.py bg$k $secs 37 bg >> gpurun_out/contention2.txt 2>&1 & done
  for k in 1 2 3 4 5 6; do env "$@" timeout 60 python scripts/gpu_contention_probe.py $name.$k $secs 37 probe $extra >> gpurun_out/contention2.txt 2>&1 & done
  wait
}
echo "== split + hand-over (QS_PDL=3), side stream waits for the default stream (fixed test)" >> gpurun_out/contention2.txt
run fixed 12 "" QS_PDL=3
echo "== control: same without the wait (the test as it was)" >> gpurun_out/contention2.txt
run control 6 --no-wait-stream QS_PDL=3
cat gpurun_out/contention2.txt
timeout 100 python -m pytest tests/test_gpu_api.py tests/test_gpu_batched.py -m gpu -q -x -k "back_to_back or one_output_array or block_chained_wrapped" 2>&1 | tail -4 | tee gpurun_out/contention2_pytest.txt
timeout 100 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 10 2>gpurun_out/contention2_bench.err | tail -1 | cut -c1-300 | tee gpurun_out/contention2_bench.json
