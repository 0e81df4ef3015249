#!/bin/bash
# round 2, GPU call 13: k_lattice_cta (TMA boxes per CTA) -- parity first, then A/B against the per-warp boxes (path 7) on the same box
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --durations=5 > gpurun_out/r2/pytest13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest13.log
tail -25 gpurun_out/r2/pytest13.log
for p in 7 9 7 9; do timeout 300 python tools/path_sweep.py 256 $p >> gpurun_out/r2/sweep13.log 2>&1; done
timeout 300 python tools/path_sweep.py 128 7 9 >> gpurun_out/r2/sweep13.log 2>&1
timeout 300 python tools/path_sweep.py robots1 4096 7 9 >> gpurun_out/r2/sweep13.log 2>&1
timeout 300 python tools/path_sweep.py holes 192 7 0 >> gpurun_out/r2/sweep13.log 2>&1
cat gpurun_out/r2/sweep13.log
