#!/bin/bash
# round 2, GPU call 15: static solve -- tests again, timing at 1 M and 6 M voxels, one ncu capture of the two iteration kernels
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_static_solve.py tests/test_dropin_cpp.py -m gpu -q --durations=5 > gpurun_out/r2/pytest15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest15.log
tail -15 gpurun_out/r2/pytest15.log
timeout 900 python tools/linsolve_bench.py 64 16 16 128 32 32 256 64 64 384 128 128 > gpurun_out/r2/linsolve15.jsonl 2> gpurun_out/r2/linsolve15.err
cat gpurun_out/r2/linsolve15.jsonl; tail -3 gpurun_out/r2/linsolve15.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lin_step -s 200 -c 2 -f -o gpurun_out/r2/linsolve_ncu \
  python tools/linsolve_bench.py 256 64 64 > gpurun_out/r2/linsolve_ncu.log 2>&1
tail -3 gpurun_out/r2/linsolve_ncu.log
