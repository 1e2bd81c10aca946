#!/bin/bash
cd "$(dirname "$0")/.."
for c in 2 4 8 16 32; do
  echo "== FGNN_TC_CHUNK=$c"
  FGNN_TC_CHUNK=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pairs/s %.1f ms/step %.2f launches %d mlp_ms %.1f matmul_ms %.1f e2e %.1f' % (d['value'], d['ms_per_step'], d['gpu_launches'], d['kernels']['tc_mlp_kernel']['total_ms']/d['steps'], d['kernels']['tc_matmul_kernel']['total_ms']/d['steps'], d['e2e']['value']))"
done
