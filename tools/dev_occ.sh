#!/bin/bash
# kernel tuning: bench one library variant with an optional shared-memory pad (occupancy control)
# usage: tools/dev_occ.sh <lib suffix or "default"> <pad bytes> [precision]
lib=$1; pad=$2; prec=${3:-f64}
if [ "$lib" != "default" ]; then export SHC_B200_LIB=$PWD/gpurun_in_lib_$lib.so; fi
export SHC_SMEM_PAD=$pad
python bench.py --steps 30 --no-cpu-baseline --precision $prec 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib pad=$pad $prec', 'value %.4g steps/s  %.1f us/step  frac %.3f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['frac']))"
