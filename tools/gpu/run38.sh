#!/bin/bash
# round 2, GPU call 38 (1 GPU): compute-sanitizer memcheck over the fused path and the new slab / ambient / Poisson-push kernels
mkdir -p gpurun_out/r2
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py 7 > gpurun_out/r2/memcheck38.log 2>&1; echo "memcheck rc=$?"
tail -14 gpurun_out/r2/memcheck38.log
