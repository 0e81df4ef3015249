#!/bin/bash
# round 2, GPU call 34 (1 GPU): slabbed checkpoints, facade on several devices again
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_slabbed.py tests/test_gpu_ambient.py -m gpu -q > gpurun_out/r2/pytest34.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest34.log
tail -12 gpurun_out/r2/pytest34.log
timeout 300 tests/cpp/_build/dropin_b200 slabbedDevices > gpurun_out/r2/dropin34.log 2>&1; echo "slabbedDevices rc=$?"; tail -4 gpurun_out/r2/dropin34.log
VX_DEVICES=0,0,0 timeout 600 tests/cpp/_build/dropin_b200 > gpurun_out/r2/dropin34_env.log 2>&1; echo "VX_DEVICES dropin rc=$?"; grep -v "^PASS" gpurun_out/r2/dropin34_env.log | tail -6
