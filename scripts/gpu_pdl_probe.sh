#!/bin/bash
# Dev probe: step time vs batch size for the PDL modes (2 = grid-wide wait + late trigger, 3 = per-block hand-over).
export OMP_NUM_THREADS=1 PYTHONUNBUFFERED=1
mkdir -p gpurun_out
out=gpurun_out/pdl_probe.txt
: > $out
run() {  # label, env assignments..., -- bench args
    label=$1; shift
    us=$(env "$@" timeout 100 python bench.py --no-extras --no-cpu-baseline --e2e-steps 10 --steps 30000 --warmup 512 $ARGS 2>/dev/null | tail -1 | python -c "import sys,json; print('%.3f' % (json.loads(sys.stdin.read())['ms_per_step']*1e3))")
    echo "$label $ARGS : $us us/step" | tee -a $out
}
for E in 32 1024 2048 4096 8192; do
  for m in 2 3; do ARGS="--config c3 --envs $E" run "QS_PDL=$m QS_SPLIT=0" QS_PDL=$m QS_SPLIT=0; done
done
ARGS="--config c3 --envs 4096" run "QS_PDL=3 QS_BLOCK=32" QS_PDL=3 QS_BLOCK=32 QS_SPLIT=0
ARGS="--config c3 --envs 4096" run "QS_PDL=3 QS_BLOCK=128" QS_PDL=3 QS_BLOCK=128 QS_SPLIT=0
ARGS="--config c3 --envs 4096" run "QS_PDL=3 QS_SPLIT=1" QS_PDL=3 QS_SPLIT=1
ARGS="--config c3 --envs 4096" run "QS_PDL=2 QS_SPLIT=1" QS_PDL=2 QS_SPLIT=1
ARGS="--config c3 --envs 4096 --groups 2" run "QS_PDL=3 groups2" QS_PDL=3
ARGS="--config c2 --envs 1024" run "QS_PDL=3 QS_SPLIT=0" QS_PDL=3 QS_SPLIT=0
ARGS="--config c2 --envs 4096" run "QS_PDL=3 QS_SPLIT=0" QS_PDL=3 QS_SPLIT=0
ARGS="--config c2 --envs 4096" run "QS_PDL=2 QS_SPLIT=0" QS_PDL=2 QS_SPLIT=0
ARGS="--config c2 --envs 4096" run "QS_PDL=3 QS_SPLIT=1" QS_PDL=3 QS_SPLIT=1
