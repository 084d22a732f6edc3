#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_batched.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2r_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2r_bench.json
python - <<'PY' | tee gpurun_out/r2r_summary.txt
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
print('c3', round(d['ms_per_step']*1e3,3), 'frac', round(d['roofline']['frac'],4), 'e2e', d['e2e']['value'])
for k in ('rollout','wrapped','large_batch','c5','configs'):
    print(k, json.dumps(d.get(k))[:600])
PY
for tool in memcheck racecheck; do echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/gpu_sanitize.py 2>&1 | tail -8; done 2>&1 | tee gpurun_out/r2r_sanitize.txt
