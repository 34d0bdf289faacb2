#!/bin/bash
# ptxas statistics (registers / spills) of the straight-line kernels of one layout, with extra -D flags:
#   tools/ptxas_probe.sh [layout index] [part] [extra nvcc flags...]
cd "$(dirname "$0")/.."
i=${1:-2}; part=${2:-0}; shift 2
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -Xptxas -v -I include -I rl4mm_b200/csrc \
  -DLOBSIM_LAYOUT_INDEX=$i -DLOBSIM_TU_PART=$part "$@" -c rl4mm_b200/csrc/fast_layout.cu -o /tmp/probe_$i_$part.o 2>&1 \
  | grep -A2 "Compiling entry" | grep -E "Compiling|registers|spill" | sed -E 's/.*(k_[a-z_]+fast)I12StaticLayoutILi([0-9]+)ELi([0-9]+)ELi([0-9]+)EE(Lb[01]ELb[01]E)?.*/\1<\2,\3,\4> \5/'
