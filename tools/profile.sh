#!/bin/bash
# Run under gpurun: launch list (share of the step per kernel) + one ncu --set full capture of the control-cycle kernel,
# plus the disassembly (with line info) of the very library that was profiled, so that tools/ncu_lines.py can join them.
# Usage: tools/profile.sh <tag> [precision] [workload]
set -x
TAG=${1:-r2}
PREC=${2:-f64}
WL=${3:-hexapod}
mkdir -p gpurun_out
LIB=${SHC_B200_LIB:-syropod_highlevel_controller_b200/libshc_b200.so}
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/launches_${TAG}_${WL}_${PREC}.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision ${PREC} --workload ${WL} > gpurun_out/bench_under_ncu_${TAG}_${WL}_${PREC}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:control_cycle -s 303 -c 2 -f -o gpurun_out/prof_${TAG}_${WL}_${PREC} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --precision ${PREC} --workload ${WL} > gpurun_out/bench_under_ncu_full_${TAG}_${WL}_${PREC}.log 2>&1
(cd gpurun_out && cuobjdump -xelf all ../$LIB > /dev/null && for f in *.cubin; do nvdisasm --print-line-info $f > sass_${TAG}.txt 2>/dev/null; rm -f $f; done; gzip -f sass_${TAG}.txt)
ls -la gpurun_out/ | tail -8
