#!/bin/bash
# round 2, GPU call 35 (1 GPU): stateInfo of slabbed models
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_slabbed.py -m gpu -q -x > gpurun_out/r2/pytest35.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest35.log
tail -12 gpurun_out/r2/pytest35.log
timeout 300 tests/cpp/_build/dropin_b200 slabbedDevices > gpurun_out/r2/dropin35.log 2>&1; echo "slabbedDevices rc=$?"; tail -4 gpurun_out/r2/dropin35.log
