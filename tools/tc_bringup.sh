#!/bin/bash
# Runs every tensor-core bring-up stage in its own process with a hard timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for st in matmul64 matmul128 matmul256 matmul_ragged mlp64 mlp32 embed embed_ragged; do
  for pr in ${PRECS:-bf16}; do
    echo "=== $st $pr"
    timeout 120 python tools/tc_check.py $st $pr 2>&1 | grep -v Warning | tail -12
  done
done
