#!/bin/bash
mkdir -p gpurun_out
export QS_LIB=$PWD/tune/libquadswarm_tl.so
timeout 300 python scripts/gpu_timeline_wrapped.py 2>&1 | tee gpurun_out/r2t_timeline_wrapped.txt | tail -8
QS_WRAP_REPLAY=0 timeout 300 python scripts/gpu_timeline_wrapped.py 2>&1 | tee gpurun_out/r2t_timeline_wrapped_noreplay.txt | tail -6
