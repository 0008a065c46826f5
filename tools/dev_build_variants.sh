#!/bin/bash
# Kernel-tuning helper: builds gpurun_in_lib_<name>.so for each "name:flags" argument in parallel (flags = nvcc -D options).
cd "$(dirname "$0")/.."
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared $flags \
    -o gpurun_in_lib_$name.so syropod_highlevel_controller_b200/csrc/shc_engine.cu 2>/dev/null &
done
wait
ls -la gpurun_in_lib_*.so
