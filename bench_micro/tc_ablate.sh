#!/bin/bash
# Ablation of prefilter_tc_kernel: one library per MSB_TC_EXP value (built here, no GPU needed),
# then on the GPU box:  bash bench_micro/tc_ablate.sh run   -> prefilter ms per variant.
set -e
cd "$(dirname "$0")/.."
if [ "$1" = "run" ]; then
  for e in ${EXPS:-0 1 2 3}; do
    MSB200_LIB=$PWD/bench_micro/libmsb200_exp$e.so python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('MSB_TC_EXP=$e', 'prefilter_ms', round(d['phase_ms']['prefilter'],3), 'step_ms', round(d['ms_per_step'],3), 'clocks', d['clocks'].get('sm_mhz'), d['clocks'].get('sm_min_mhz'), 'power_w_max', d['clocks'].get('power_w_max'), d['clocks'].get('reasons'))"
  done
else
  for e in ${EXPS:-0 1 2 3}; do
    python -c "from motifscan_b200 import build; build.build(out='bench_micro/libmsb200_exp$e.so', defines=('MSB_TC_EXP=$e',))"
  done
fi
