#!/bin/bash
# round 2, GPU call 30 (1 GPU): one 6x6x2 pose box per brick instead of own + four face boxes (9 TMA copies instead of 13): parity, then A/B timing on one box
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabbed.py -m gpu -q -x > gpurun_out/r2/pytest30.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest30.log
tail -6 gpurun_out/r2/pytest30.log
for rep in 1 2; do
for v in base xy; do
  echo "== $v" >> gpurun_out/r2/ab30.log
  VX_PRODUCT_SO=$PWD/voxelyze_b200/lib/variants/$v.so timeout 600 python tools/path_sweep.py 256 0 >> gpurun_out/r2/ab30.log 2>&1
done; done
VX_PRODUCT_SO=$PWD/voxelyze_b200/lib/variants/base.so timeout 600 python tools/path_sweep.py robots 4096 0 >> gpurun_out/r2/ab30.log 2>&1
VX_PRODUCT_SO=$PWD/voxelyze_b200/lib/variants/xy.so timeout 600 python tools/path_sweep.py robots 4096 0 >> gpurun_out/r2/ab30.log 2>&1
cat gpurun_out/r2/ab30.log
