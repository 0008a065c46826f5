#!/bin/bash
# Run under gpurun: launch list (share of the step per kernel) + one ncu --set full capture of the control-cycle kernel.
# Usage: tools/profile.sh <tag> [precision]
set -x
TAG=${1:-r1}
PREC=${2:-f64}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/launches_${TAG}_${PREC}.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision ${PREC} > gpurun_out/bench_under_ncu_${TAG}_${PREC}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:control_cycle -s 303 -c 2 -f -o gpurun_out/prof_${TAG}_${PREC} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --precision ${PREC} > gpurun_out/bench_under_ncu_full_${TAG}_${PREC}.log 2>&1
ls -la gpurun_out/
