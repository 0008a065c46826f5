#!/bin/bash
# Under gpurun --gpus N: the fused gather with its landed signal issued every 1 / 4 / 8 cycles, by a kernel or by a stream
# memory operation, and without signals (tuning reference).  Usage: NG=2 TAG=r2p tools/dev_signal_variants.sh
NG=${NG:-2}; TAG=${TAG:-sig}
mkdir -p gpurun_out
source tools/dev_multi_bench.sh
run k1 "SHC_GATHER_SIGNAL_EVERY=1" --no-cpu-baseline
run k4 "SHC_GATHER_SIGNAL_EVERY=4" --no-cpu-baseline
run k8 "SHC_GATHER_SIGNAL_EVERY=8" --no-cpu-baseline
run m1 "SHC_GATHER_SIGNAL=memop SHC_GATHER_SIGNAL_EVERY=1" --no-cpu-baseline
run m4 "SHC_GATHER_SIGNAL=memop SHC_GATHER_SIGNAL_EVERY=4" --no-cpu-baseline
run none "SHC_GATHER_TUNE=nosignal" --no-cpu-baseline --no-verify-gather
