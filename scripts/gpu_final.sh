#!/bin/bash
# Round-end verification on the GPU box: full GPU test-suite, smoke, sanitizer, the driver's bench commands.
export OMP_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 MKL_NUM_THREADS=1 PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -6
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
: > gpurun_out/sanitizer.log
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool python scripts/gpu_sanitize.py" >> gpurun_out/sanitizer.log
  timeout 600 compute-sanitizer --tool $tool python scripts/gpu_sanitize.py 2>&1 | grep -v "^=========     \(Host Frame\|Saved host\)" | tail -30 >> gpurun_out/sanitizer.log
done
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|===" gpurun_out/sanitizer.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/bench_final.err | tail -1 > gpurun_out/bench_final.json
cut -c1-400 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_reference_arm.json; cut -c1-400 gpurun_out/bench_reference_arm.json
