#!/bin/bash
# round 2, GPU call 45 (1 GPU): slabbedDevices of the drop-in source (device moves with a pending edit, checkpoint of the edited model)
mkdir -p gpurun_out/r2
timeout 60 tests/cpp/_build/dropin_b200 slabbedDevices > gpurun_out/r2/dropin45.log 2>&1; echo "slabbedDevices rc=$?"; tail -3 gpurun_out/r2/dropin45.log
