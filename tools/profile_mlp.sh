#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-x}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tc_mlp_kernel -s 4 -c 2 -f -o gpurun_out/prof_mlp_${TAG} \
    python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
