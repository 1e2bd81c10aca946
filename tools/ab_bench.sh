#!/bin/bash
# A/B the libraries in tmp_variants/ on ONE box: alternate them so that box-to-box variance (clocks, power cap) cancels.
# usage: tools/ab_bench.sh VARIANT...   (tmp_variants/lib<VARIANT>.so; "name:ENV=1" sets an environment variable)
cd "$(dirname "$0")/.."
cp graph_neural_net_b200/csrc/libfgnn_b200.so /tmp/lib_orig.so
for rep in 1 2 3; do
  for spec in "$@"; do
    v=${spec%%:*}; e=""; [[ "$spec" == *:* ]] && e=${spec#*:}
    cp tmp_variants/lib$v.so graph_neural_net_b200/csrc/libfgnn_b200.so
    r=$(env $e timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary 2>&1 | grep -o '"value": [0-9.]*\|tc_m[a-z]*_kernel": {"launches": [0-9]*, "total_ms": [0-9.]*' | head -3 | tr '\n' ' ')
    echo "rep $rep variant $spec $r"
  done
done
cp /tmp/lib_orig.so graph_neural_net_b200/csrc/libfgnn_b200.so
