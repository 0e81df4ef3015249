#!/bin/bash
# round 2, GPU call 40 (1 GPU): slab tests after the exchange counters restart on detach
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_slabbed.py -m gpu -q > gpurun_out/r2/pytest40.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest40.log
tail -12 gpurun_out/r2/pytest40.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "slab or split or peer or halo" > gpurun_out/r2/pytest40b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest40b.log
tail -3 gpurun_out/r2/pytest40b.log
