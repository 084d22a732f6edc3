#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_profile.sh > gpurun_out/profile.log 2>&1
export QS_LIB=$PWD/tune/libquadswarm_tl.so
timeout 200 python scripts/gpu_timeline.py c3 stagger 2>&1 | tee gpurun_out/r2x_timeline_c3.txt | tail -12
timeout 300 python scripts/gpu_timeline_wrapped.py 2>&1 | tee gpurun_out/r2x_timeline_wrapped.txt | tail -6
unset QS_LIB
python -m pytest tests/test_gpu_api.py tests/test_gpu_batched.py -m gpu -q -x 2>&1 | tail -3
