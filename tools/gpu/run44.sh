#!/bin/bash
# round 2, GPU call 44 (1 GPU): the facade's device moves with a pending edit
mkdir -p gpurun_out/r2
timeout 60 tests/cpp/_build/dropin_b200 slabbedDevices > gpurun_out/r2/dropin44.log 2>&1; echo "slabbedDevices rc=$?"; tail -4 gpurun_out/r2/dropin44.log
VX_DEVICES=0,0,0 timeout 100 tests/cpp/_build/dropin_b200 > gpurun_out/r2/dropin44_env.log 2>&1; echo "VX_DEVICES dropin rc=$?"; grep -v "^PASS" gpurun_out/r2/dropin44_env.log | tail -4
