#!/bin/bash
# round 2, GPU call 7: full GPU suite (Poisson on the fused path, slabs from scenarios), A/B base vs current kernel, C3 over 3000 steps
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2/pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest7.log
tail -30 gpurun_out/r2/pytest7.log
for v in base main base main; do
  if [ $v = main ]; then unset VX_PRODUCT_SO; else export VX_PRODUCT_SO=$PWD/voxelyze_b200/lib/variants/lib$v.so; fi
  echo "== $v" >> gpurun_out/r2/sweep7.log
  timeout 300 python tools/path_sweep.py 256 0 >> gpurun_out/r2/sweep7.log 2>&1
done
unset VX_PRODUCT_SO
cat gpurun_out/r2/sweep7.log
timeout 900 python tools/config_bench.py --config c3 --steps 3000 --warmup 100 > gpurun_out/r2/configs7.jsonl 2> gpurun_out/r2/configs7.err
cat gpurun_out/r2/configs7.jsonl; tail -3 gpurun_out/r2/configs7.err
