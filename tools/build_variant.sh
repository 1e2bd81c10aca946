#!/bin/bash
# Build a variant of libfgnn_b200.so out of tree: tools/build_variant.sh NAME [EXTRA nvcc flags...] -> tmp_variants/libNAME.so
# (sources are copied from the working tree as they are now; used with tools/ab_bench.sh to A/B builds on one box)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
D=/tmp/fgnn_variant_$NAME
rm -rf $D; mkdir -p $D/graph_neural_net_b200 $D/include tmp_variants
cp -r graph_neural_net_b200/csrc $D/graph_neural_net_b200/
cp include/*.h $D/include/
rm -f $D/graph_neural_net_b200/csrc/*.o $D/graph_neural_net_b200/csrc/*.so
make -C $D/graph_neural_net_b200/csrc -j8 EXTRA="$*" > $D/build.log 2>&1 || { tail -30 $D/build.log; exit 1; }
cp $D/graph_neural_net_b200/csrc/libfgnn_b200.so tmp_variants/lib$NAME.so
echo "built tmp_variants/lib$NAME.so"
