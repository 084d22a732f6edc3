#!/bin/bash
mkdir -p gpurun_out
export QS_LIB=$PWD/tune/libquadswarm_tl.so
for pdl in 2 3; do QS_PDL=$pdl timeout 200 python scripts/gpu_timeline.py c3 stagger; QS_PDL=$pdl timeout 200 python scripts/gpu_timeline.py c3; done 2>&1 | tee gpurun_out/r2k_timeline.txt
unset QS_LIB
for tool in memcheck racecheck synccheck; do echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/gpu_sanitize.py 2>&1 | tail -25; done 2>&1 | tee gpurun_out/r2k_sanitize.txt
