#!/bin/bash
# kernel tuning: octopod (configs[3]) bench of one library variant
lib=$1; prec=${2:-f64}
if [ "$lib" != "default" ]; then export SHC_B200_LIB=$PWD/gpurun_in_lib_$lib.so; fi
python bench.py --workload octopod --steps 15 --warmup 3 --no-cpu-baseline --precision $prec 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib $prec octopod', 'value %.4g steps/s  %.1f us/step  frac %.3f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['frac']))"
