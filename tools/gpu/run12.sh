#!/bin/bash
# round 2, GPU call 12: A/B of shared-memory material tables on C4 and C3, full GPU suite
mkdir -p gpurun_out/r2
for v in base main base main; do
  if [ $v = main ]; then unset VX_PRODUCT_SO; else export VX_PRODUCT_SO=$PWD/voxelyze_b200/lib/variants/lib$v.so; fi
  echo "== $v" >> gpurun_out/r2/sweep12.log
  timeout 300 python bench.py --config c4 --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4', d['ms_per_step'], d['roofline']['launch_ms'])" >> gpurun_out/r2/sweep12.log 2>&1
  timeout 300 python tools/config_bench.py --config c3 --steps 1000 --warmup 100 --c3-warmup 200 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['ms_per_step'])" >> gpurun_out/r2/sweep12.log 2>&1
done
unset VX_PRODUCT_SO
cat gpurun_out/r2/sweep12.log
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2/pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest12.log
tail -12 gpurun_out/r2/pytest12.log
