#!/bin/bash
# round 2, GPU call 10: full GPU suite (sparse bodies, per-voxel floor), ncu capture of the fused kernel on the C4 workload
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2/pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest10.log
tail -25 gpurun_out/r2/pytest10.log
bash tools/ncu_c4.sh r2/c4a
