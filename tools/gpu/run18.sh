#!/bin/bash
# round 2, GPU call 18: where k_small_steps spends a step (C1): phases switched off one at a time (timing only, states are meaningless)
mkdir -p gpurun_out/r2
for v in main sp1 sp2 sp4 sp3 main; do
  if [ $v = main ]; then unset VX_PRODUCT_SO; else export VX_PRODUCT_SO=$PWD/voxelyze_b200/lib/variants/lib$v.so; fi
  echo "== $v" >> gpurun_out/r2/small18.log
  timeout 300 python tools/config_bench.py --config c1 --steps 10000 --warmup 200 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['config'], d['kernel'][:24], round(d['ms_per_step'] * 1e3, 3), 'us/step')" >> gpurun_out/r2/small18.log
done
cat gpurun_out/r2/small18.log
