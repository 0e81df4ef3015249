#!/bin/bash
# round 2, GPU call 20: full GPU suite with the final small-model launch shape, smoke, C1 timing
mkdir -p gpurun_out/r2
timeout 2400 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2/pytest20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest20.log
tail -14 gpurun_out/r2/pytest20.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python tools/config_bench.py --config c1 --steps 10000 --warmup 200 2>/dev/null | tee gpurun_out/r2/c1_20.jsonl
VX_NO_SMALL=1 timeout 300 python tools/config_bench.py --config c1 --steps 10000 --warmup 200 2>/dev/null | tee -a gpurun_out/r2/c1_20.jsonl
