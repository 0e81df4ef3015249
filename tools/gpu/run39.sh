#!/bin/bash
# round 2, GPU call 39 (1 GPU): Poisson's ratio changed (non-zero to non-zero) mid-run on a slabbed model
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_slabbed.py tests/test_gpu_ambient.py -m gpu -q > gpurun_out/r2/pytest39.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest39.log
tail -12 gpurun_out/r2/pytest39.log | cut -c1-300
