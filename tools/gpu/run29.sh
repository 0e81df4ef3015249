#!/bin/bash
# round 2, GPU call 29 (1 GPU): Poisson materials across slabs (pStrain in the peer stores)
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_slabbed.py -m gpu -q -x > gpurun_out/r2/pytest29a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest29a.log
tail -30 gpurun_out/r2/pytest29a.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "poisson or slab or split or peer or halo or bitwise" > gpurun_out/r2/pytest29b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest29b.log
tail -5 gpurun_out/r2/pytest29b.log
