#!/bin/bash
# Under gpurun (1 GPU): the bench lines of BASELINE.md section 4 — configs 2, 3 (gait sweep), 4 and 5's single-GPU shard, plus the
# reference arm.  Usage: TAG=r2u tools/baseline_table.sh
TAG=${TAG:-r2u}
mkdir -p gpurun_out
b() { name=$1; shift; python bench.py "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; tail -c 300 gpurun_out/${TAG}_$name.err; }
b config2_4096_tripod --robots-per-gpu 4096 --steps 200 --warmup 20 --no-cpu-baseline
for g in wave amble ripple tripod; do
  b config3_65536_${g} --robots-per-gpu 65536 --gait ${g}_gait --steps 50 --warmup 5 --no-cpu-baseline
done
b config3_65536_tripod_mixed --robots-per-gpu 65536 --gait tripod_gait --precision mixed --steps 50 --warmup 5 --no-cpu-baseline
b config4_octopod --workload octopod --steps 20 --warmup 3 --no-cpu-baseline
b config4_octopod_mixed --workload octopod --precision mixed --steps 20 --warmup 3 --no-cpu-baseline
b config5_shard_f64 --steps 50 --warmup 5
b config5_shard_mixed --precision mixed --steps 50 --warmup 5 --no-cpu-baseline
b reference_arm --impl reference --steps 3 --warmup 1
python tools/baseline_rows.py gpurun_out/${TAG}_*.json
