#!/bin/bash
# round 2, GPU call 6: full GPU suite (mesh, gtests, Poisson toggle), C3 with live contacts, facade-free c4/c5 lines
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2/pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest6.log
tail -30 gpurun_out/r2/pytest6.log
timeout 900 python tools/config_bench.py --config c3 --steps 1000 --warmup 100 > gpurun_out/r2/configs6.jsonl 2> gpurun_out/r2/configs6.err
cat gpurun_out/r2/configs6.jsonl; tail -3 gpurun_out/r2/configs6.err
