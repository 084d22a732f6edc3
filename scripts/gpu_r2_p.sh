#!/bin/bash
mkdir -p gpurun_out
QS_SPLIT=0 QS_BALANCE_KB=48 timeout 200 python bench.py --config c2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 10 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2p_err.txt
QS_SPLIT=0 QS_BALANCE_KB=56 timeout 200 python bench.py --config c2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 10 2>&1 | tail -2 | cut -c1-200 | tee -a gpurun_out/r2p_err.txt
