#!/bin/bash
# round 2, GPU call 37 (1 GPU): final -- the whole GPU suite, smoke(), the bench line of the final build
mkdir -p gpurun_out/r2
timeout 2400 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2/pytest37.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest37.log
tail -14 gpurun_out/r2/pytest37.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke37.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2/smoke37.log
timeout 900 python bench.py --gpus 1 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2/bench37_c5_n1.json 2> gpurun_out/r2/bench37_c5_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2/bench37_c5_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','parity')}, d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('dram_frac'), d['clocks'].get('reasons'))
PY
tail -3 gpurun_out/r2/bench37_c5_n1.err
