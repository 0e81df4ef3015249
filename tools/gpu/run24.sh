#!/bin/bash
# round 2, GPU call 24 (1 GPU): vx_slabbed_* (one process, several slabs on one device: lock-step peer stores) + the slab tests after the vx_slab_step split
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_slabbed.py -m gpu -q -x --durations=5 > gpurun_out/r2/pytest24a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest24a.log
tail -30 gpurun_out/r2/pytest24a.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slab or split or peer or halo" > gpurun_out/r2/pytest24b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest24b.log
tail -5 gpurun_out/r2/pytest24b.log
