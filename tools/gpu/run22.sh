#!/bin/bash
# round 2, GPU call 22 (1 GPU): z-slabs without ghost-plane bricks (GSKIP) -- the one-GPU slab emulation tests, then the whole suite
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slab or split or peer or halo" --durations=5 > gpurun_out/r2/pytest22a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest22a.log
tail -25 gpurun_out/r2/pytest22a.log
timeout 2400 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2/pytest22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest22.log
tail -12 gpurun_out/r2/pytest22.log
