#!/bin/bash
# Round-end verification on the GPU box: full GPU test-suite, smoke, sanitizer, ncu passes, default bench.
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1 PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -6
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
: > gpurun_out/sanitizer.log
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== compute-sanitizer --tool $tool python scripts/gpu_sanitize.py" >> gpurun_out/sanitizer.log
  timeout 420 compute-sanitizer --tool $tool python scripts/gpu_sanitize.py 2>&1 | grep -v "^=========     \(Host Frame\|Saved host\)" | tail -40 >> gpurun_out/sanitizer.log
done
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|===" gpurun_out/sanitizer.log
bash scripts/gpu_profile.sh > gpurun_out/profile.log 2>&1
timeout 300 python bench.py 2>gpurun_out/bench_final.err | tail -1 > gpurun_out/bench_final.json
cut -c1-250 gpurun_out/bench_final.json
for c in c2 c4; do timeout 200 python bench.py --config $c --no-cpu-baseline --e2e-steps 50 2>/dev/null | tail -1 > gpurun_out/bench_final_$c.json; cut -c1-160 gpurun_out/bench_final_$c.json; done
timeout 400 python bench.py --impl reference --steps 300 --warmup 30 2>/dev/null | tail -1 > gpurun_out/bench_reference_arm.json; cut -c1-300 gpurun_out/bench_reference_arm.json
