#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --config c3 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 10 --wrapped-main 2>&1 | tail -12 | cut -c1-400 | tee gpurun_out/r2p_err.txt
