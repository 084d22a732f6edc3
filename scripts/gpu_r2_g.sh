#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2g_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -3 gpurun_out/r2g_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
print('us/step', d['ms_per_step'] * 1e3, 'frac', d['roofline']['frac'])
print({k: v for k, v in d['e2e'].items() if k != 'note'})
print('wrapped', d['wrapped']['us_per_step'], 'cpu', d['cpu_baseline'])
PY
