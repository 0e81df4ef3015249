#!/bin/bash
# round 2, GPU call 41 (1 GPU): A/B of cache-hint variants of the fused kernel (streaming stores, evict-first record loads, L2 promotion 256 B)
mkdir -p gpurun_out/r2
timeout 200 python tools/variant_sweep.py 256 base cs ef csef cs256 > gpurun_out/r2/ab41.log 2>&1; echo "rc=$?"
cat gpurun_out/r2/ab41.log | tail -12
