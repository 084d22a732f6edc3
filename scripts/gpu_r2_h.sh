#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_batched.py -q 2>&1 | tail -25 | tee gpurun_out/r2h_pytest.txt
python -m pytest tests/test_gpu_parity.py -q -s -k "masks_bit_exact" 2>&1 | grep -v "^$" | tail -25 | tee -a gpurun_out/r2h_pytest.txt
python -m pytest tests/test_gpu_api.py -q -k "host_buffer" 2>&1 | tail -3 | tee -a gpurun_out/r2h_pytest.txt
for zb in 0 1; do echo "== QS_ZC_BULK=$zb"; QS_ZC_BULK=$zb timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']; print({k: (round(v/1e6,1) if isinstance(v,float) and v>1e6 else v) for k,v in e.items() if k not in ('note',)})"; done 2>&1 | tee gpurun_out/r2h_e2e.txt
