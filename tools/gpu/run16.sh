#!/bin/bash
# round 2, GPU call 16: k_lin_step_a variants (occupancy / hoisted neighbour loads), same box
mkdir -p gpurun_out/r2
for v in main lin_mb4 lin_mb5 lin_h1 lin_h4 main lin_mb4; do
  if [ $v = main ]; then unset VX_PRODUCT_SO; else export VX_PRODUCT_SO=$PWD/voxelyze_b200/lib/variants/lib$v.so; fi
  echo "== $v" >> gpurun_out/r2/linsolve16.log
  timeout 300 python tools/linsolve_bench.py 256 64 64 384 128 128 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['workload'], d['iterations'], round(d['ms_per_iteration'], 4), 'ms/it', round(d['gbs']), 'GB/s')
    else: print(l.rstrip())" >> gpurun_out/r2/linsolve16.log
done
cat gpurun_out/r2/linsolve16.log
