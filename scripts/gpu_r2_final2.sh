#!/bin/bash
# Last validation of the round-2 tree (10 GPU-minutes left): the device-side run_away scenario and the full-size sampled
# parity first (serial), then the whole GPU suite on 6 xdist workers, a serial re-run of anything that failed there
# (workers share the GPU by time-slicing), then the driver's bench command.
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1 PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T0=$SECONDS
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "run_away" 2>&1 | tail -15 | tee gpurun_out/f2_run_away.txt
echo "t=$((SECONDS-T0))"
timeout 330 python -m pytest tests -m gpu -q -n 6 --durations=6 2>&1 | tail -40 | tee gpurun_out/f2_pytest_xdist.txt
echo "t=$((SECONDS-T0))"
if ! grep -q " passed" gpurun_out/f2_pytest_xdist.txt || grep -q "failed\|error" gpurun_out/f2_pytest_xdist.txt; then
  timeout 200 python -m pytest tests -m gpu -q --lf -x 2>&1 | tail -30 | tee gpurun_out/f2_pytest_lf.txt
fi
echo "t=$((SECONDS-T0))"
timeout 150 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/f2_bench.err | tail -1 > gpurun_out/f2_bench.json
cut -c1-600 gpurun_out/f2_bench.json
echo "t=$((SECONDS-T0))"
