#!/bin/bash
# round 2, GPU call 42 (1 GPU): final build (streaming result stores on lattices beyond the L2, 256 B L2 promotion): whole GPU suite, bench line
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2/pytest42.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest42.log
tail -5 gpurun_out/r2/pytest42.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 --no-cpu-baseline --no-facade > gpurun_out/r2/bench42_c5_n1.json 2> gpurun_out/r2/bench42_c5_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2/bench42_c5_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('launch_ms'), d['clocks'].get('reasons'))
PY
tail -2 gpurun_out/r2/bench42_c5_n1.err
