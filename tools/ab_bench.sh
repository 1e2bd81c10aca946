#!/bin/bash
# A/B the libraries in tmp_variants/ on ONE box: alternate them so that box-to-box variance (clocks, power cap) cancels.
cd "$(dirname "$0")/.."
cp graph_neural_net_b200/csrc/libfgnn_b200.so /tmp/lib_orig.so
for rep in 1 2 3; do
  for v in "$@"; do
    cp tmp_variants/lib$v.so graph_neural_net_b200/csrc/libfgnn_b200.so
    r=$(timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"value": [0-9.]*' | head -1)
    echo "rep $rep variant $v $r"
  done
done
cp /tmp/lib_orig.so graph_neural_net_b200/csrc/libfgnn_b200.so
