#!/bin/bash
# round 2, GPU call 11: A/B of the half-angle large-angle alignment on the C4 workload, full GPU suite
mkdir -p gpurun_out/r2
for v in base main base main; do
  if [ $v = main ]; then unset VX_PRODUCT_SO; else export VX_PRODUCT_SO=$PWD/voxelyze_b200/lib/variants/lib$v.so; fi
  echo "== $v" >> gpurun_out/r2/sweep11.log
  timeout 300 python bench.py --config c4 --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['launch_ms'])" >> gpurun_out/r2/sweep11.log 2>&1
done
unset VX_PRODUCT_SO
cat gpurun_out/r2/sweep11.log
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2/pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest11.log
tail -25 gpurun_out/r2/pytest11.log
