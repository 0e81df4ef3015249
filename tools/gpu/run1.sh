#!/bin/bash
# round 2, GPU call 1: parity tests, variant sweep, bench line, ncu capture of the fused kernel
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2/smi1.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest1.log
tail -5 gpurun_out/r2/pytest1.log
for v in main noddiv; do
  if [ $v = main ]; then unset VX_PRODUCT_SO; else export VX_PRODUCT_SO=$PWD/voxelyze_b200/lib/variants/lib$v.so; fi
  echo "== $v" >> gpurun_out/r2/sweep1.log
  timeout 300 python tools/path_sweep.py 256 0 5 >> gpurun_out/r2/sweep1.log 2>&1
done
unset VX_PRODUCT_SO
cat gpurun_out/r2/sweep1.log
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2/bench1.json 2> gpurun_out/r2/bench1.err
cat gpurun_out/r2/bench1.json
bash tools/ncu_one.sh 0 k_lattice_tma r2/tma1
