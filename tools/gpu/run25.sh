#!/bin/bash
# round 2, GPU call 25 (2 GPUs): vx_slabbed_* with one slab per device (all devices queued before any is waited for), the two-rank slab tests,
# and the single-process entry timed on 2 x 256^3
mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/smi25.txt
timeout 600 python -m pytest tests/test_gpu_slabbed.py -m gpu -q -x --durations=5 > gpurun_out/r2/pytest25a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest25a.log
tail -12 gpurun_out/r2/pytest25a.log
timeout 600 python tools/slabbed_bench.py 256 256 512 2 200 > gpurun_out/r2/slabbed25_n2.json 2> gpurun_out/r2/slabbed25_n2.err; echo "slabbed n2 rc=$?"; cat gpurun_out/r2/slabbed25_n2.json; tail -3 gpurun_out/r2/slabbed25_n2.err
timeout 600 python tools/slabbed_bench.py 256 256 256 1 200 > gpurun_out/r2/slabbed25_n1.json 2> gpurun_out/r2/slabbed25_n1.err; echo "slabbed n1 rc=$?"; cat gpurun_out/r2/slabbed25_n1.json; tail -3 gpurun_out/r2/slabbed25_n1.err
