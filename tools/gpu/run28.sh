#!/bin/bash
# round 2, GPU call 28 (1 GPU): whole GPU suite after vx_slabbed / setDevices / vx_set_clock
mkdir -p gpurun_out/r2
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2/pytest28.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest28.log
tail -16 gpurun_out/r2/pytest28.log
