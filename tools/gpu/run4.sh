#!/bin/bash
# round 2, GPU call 4: full GPU suite (no -x), c4 bench, path sweep for ensembles, C3 full-size with contacts
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2/pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest4.log
tail -25 gpurun_out/r2/pytest4.log
timeout 900 python bench.py --config c4 --steps 200 --warmup 20 > gpurun_out/r2/bench4_c4.json 2> gpurun_out/r2/bench4_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2/bench4_c4.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['roofline']['kernel'], d['roofline']['launch_ms'], d['roofline']['frac']); print(d.get('cpu_baseline'))
except Exception as e: print("c4 parse", e)
PY
tail -3 gpurun_out/r2/bench4_c4.err
VX_NO_PACK=1 timeout 300 python tools/path_sweep.py robots 4096 0 5 7 > gpurun_out/r2/sweep4.log 2>&1
timeout 300 python tools/path_sweep.py robots 4096 0 5 7 >> gpurun_out/r2/sweep4.log 2>&1
cat gpurun_out/r2/sweep4.log
timeout 900 python tools/config_bench.py --config c1,c2,c3 --steps 1000 --warmup 100 > gpurun_out/r2/configs4.jsonl 2> gpurun_out/r2/configs4.err
cat gpurun_out/r2/configs4.jsonl; tail -3 gpurun_out/r2/configs4.err
