#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_batched.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2s_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2s_bench.json
python - <<'PY' | tee gpurun_out/r2s_summary.txt
import json
d=json.load(open('gpurun_out/r2s_bench.json'))
print('c3', round(d['ms_per_step']*1e3,3), 'frac', round(d['roofline']['frac'],4), 'e2e', d['e2e']['value'])
print('wrapped', json.dumps(d.get('wrapped'))[:300])
PY
