#!/bin/bash
# round 2, GPU call 26 (1 GPU): the facade on several devices (setDevices, VX_DEVICES)
mkdir -p gpurun_out/r2
timeout 300 tests/cpp/_build/dropin_b200 slabbedDevices > gpurun_out/r2/dropin26.log 2>&1; echo "slabbedDevices rc=$?"; tail -5 gpurun_out/r2/dropin26.log
VX_DEVICES=0,0,0 timeout 600 tests/cpp/_build/dropin_b200 > gpurun_out/r2/dropin26_env.log 2>&1; echo "VX_DEVICES dropin rc=$?"; grep -v "^PASS" gpurun_out/r2/dropin26_env.log | tail -15
timeout 900 python -m pytest tests/test_dropin_cpp.py -m gpu -q -x > gpurun_out/r2/pytest26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest26.log
tail -8 gpurun_out/r2/pytest26.log
