#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of the tcgen05 kernels.
# usage (under gpurun): bash tools/profile_round.sh <tag>
cd "$(dirname "$0")/.."
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-secondary"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# conv chain: the mlp1|mlp2 launch (NMLP=2), the mlp3 launch (NMLP=1) and the pooled one
ncu --set full --clock-control none --import-source on -k regex:tc_mlp_kernel -s 6 -c 3 -f -o gpurun_out/prof_mlp_${TAG} $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_matmul_kernel -s 2 -c 1 -f -o gpurun_out/prof_matmul_${TAG} $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_head_kernel -s 1 -c 1 -f -o gpurun_out/prof_head_${TAG} $B > /dev/null 2>&1
ls -la gpurun_out/
