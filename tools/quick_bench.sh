#!/bin/bash
# quick device-side check under gpurun: parity smoke + bench value for both precisions
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for p in f64 mixed; do
python bench.py --steps 30 --no-cpu-baseline --precision $p 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$p', 'value %.4g steps/s  %.1f us/step  frac %.3f  e2e %.4g' % (d['value'], d['ms_per_step']*1e3, d['roofline']['frac'], d['e2e']['value']))"
done
