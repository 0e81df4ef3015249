#!/bin/bash
# round 2, GPU call 32 (2 GPUs): slab tests with one slab per device after the Poisson / ambient / split work
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_slabbed.py tests/test_gpu_slab.py -m gpu -q > gpurun_out/r2/pytest32.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest32.log
tail -8 gpurun_out/r2/pytest32.log
