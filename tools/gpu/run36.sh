#!/bin/bash
# round 2, GPU call 36 (2 GPUs): vx_slabbed with a host thread per device (build and stepping), against VX_SLABBED_THREADS=0
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_slabbed.py -m gpu -q -k "per_device or checkpoint or peer_stores" > gpurun_out/r2/pytest36.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest36.log
tail -4 gpurun_out/r2/pytest36.log
timeout 300 python tools/slabbed_bench.py 256 256 512 2 200 > gpurun_out/r2/slabbed36_threads.json 2> gpurun_out/r2/slabbed36_threads.err; echo "threads rc=$?"; cat gpurun_out/r2/slabbed36_threads.json; tail -2 gpurun_out/r2/slabbed36_threads.err
VX_SLABBED_THREADS=0 timeout 300 python tools/slabbed_bench.py 256 256 512 2 200 > gpurun_out/r2/slabbed36_serial.json 2> gpurun_out/r2/slabbed36_serial.err; echo "serial rc=$?"; cat gpurun_out/r2/slabbed36_serial.json; tail -2 gpurun_out/r2/slabbed36_serial.err
