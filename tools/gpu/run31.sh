#!/bin/bash
# round 2, GPU call 31 (1 GPU): vx_step_ambient (temperature program in one call) -- parity with the per-step calls, then C4 through bench.py
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_ambient.py -m gpu -q -x > gpurun_out/r2/pytest31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest31.log
tail -25 gpurun_out/r2/pytest31.log
timeout 900 python bench.py --gpus 1 --config c4 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2/bench31_c4_n1.json 2> gpurun_out/r2/bench31_c4_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2/bench31_c4_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('launch_ms'))
PY
tail -3 gpurun_out/r2/bench31_c4_n1.err
