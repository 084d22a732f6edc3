#!/bin/bash
# On-GPU check used during development: sanity probe, GPU tests under a hard timeout with streaming logs, short bench.
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1 PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 150 python scripts/gpu_sanity.py 2>&1 | tee gpurun_out/sanity.log
timeout ${TEST_TIMEOUT:-600} python -m pytest tests -m gpu -q -p no:cacheprovider -x "$@" 2>&1 | tee gpurun_out/pytest_gpu.log | tail -60
