#!/bin/bash
# round 2, GPU call 27 (1 GPU): the facade on several devices (setDevices, VX_DEVICES) with the clock carried across handles
mkdir -p gpurun_out/r2
timeout 300 tests/cpp/_build/dropin_b200 slabbedDevices > gpurun_out/r2/dropin27.log 2>&1; echo "slabbedDevices rc=$?"; tail -5 gpurun_out/r2/dropin27.log
VX_DEVICES=0,0,0 timeout 600 tests/cpp/_build/dropin_b200 > gpurun_out/r2/dropin27_env.log 2>&1; echo "VX_DEVICES dropin rc=$?"; grep -v "^PASS" gpurun_out/r2/dropin27_env.log | tail -15
VX_DEVICES=0,0 timeout 600 tests/cpp/_build/ref_gtests_b200 > gpurun_out/r2/gtests27_env.log 2>&1; echo "VX_DEVICES gtests rc=$?"; tail -4 gpurun_out/r2/gtests27_env.log
timeout 900 python -m pytest tests/test_dropin_cpp.py tests/test_gpu_slabbed.py -m gpu -q > gpurun_out/r2/pytest27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest27.log
tail -8 gpurun_out/r2/pytest27.log
