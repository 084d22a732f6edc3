#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --config c3 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 10 2>&1 | tail -30 | tee gpurun_out/r2p_err.txt
python -m pytest tests/test_gpu_api.py -m gpu -q -x -k "back_to_back" 2>&1 | tail -40 | tee -a gpurun_out/r2p_err.txt
