#!/bin/bash
# round 2, GPU call 14: static solve (vx_linear_solve) -- parity tests, drop-in C++ test, timing
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_static_solve.py tests/test_dropin_cpp.py tests/test_abi.py -m gpu -q --durations=8 > gpurun_out/r2/pytest14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest14.log
tail -40 gpurun_out/r2/pytest14.log
timeout 900 python tools/linsolve_bench.py 64 16 16 128 32 32 > gpurun_out/r2/linsolve14.jsonl 2> gpurun_out/r2/linsolve14.err
cat gpurun_out/r2/linsolve14.jsonl; tail -3 gpurun_out/r2/linsolve14.err
