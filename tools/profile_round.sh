#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of the two tcgen05 kernels.
# usage (under gpurun): bash tools/profile_round.sh <tag>
cd "$(dirname "$0")/.."
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_mlp_kernel -s 6 -c 2 -f -o gpurun_out/prof_mlp_${TAG} \
    python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_matmul_kernel -s 2 -c 2 -f -o gpurun_out/prof_matmul_${TAG} \
    python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/
