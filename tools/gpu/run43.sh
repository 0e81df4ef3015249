#!/bin/bash
# round 2, GPU call 43 (1 GPU): the slabbed tests once more on the final library (stateInfo min/max now compared exactly)
mkdir -p gpurun_out/r2
timeout 120 python -m pytest tests/test_gpu_slabbed.py -m gpu -q -x > gpurun_out/r2/pytest43.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest43.log
tail -6 gpurun_out/r2/pytest43.log | cut -c1-400
