#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_batched.py -q -x 2>&1 | tail -5 | tee gpurun_out/r2f_pytest.txt
probe() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('bare', round(d['ms_per_step']*1e3,2), 'wrapped', round(d['wrapped']['us_per_step'],2), d['wrapped']['checkpoints'], d['wrapped']['events_replayed'])"; }
probe QS_WRAP_PROBE=0 2>&1 | tee gpurun_out/r2f_probe.txt
probe QS_WRAP_REPLAY=0 2>&1 | tee -a gpurun_out/r2f_probe.txt
