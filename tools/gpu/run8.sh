#!/bin/bash
# round 2, GPU call 8: full GPU suite, bench c5 (CPU parity + facade leg), ncu launch list and full capture of the fused kernel
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2/pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest8.log
tail -12 gpurun_out/r2/pytest8.log
timeout 1500 python bench.py --steps 200 --warmup 20 > gpurun_out/r2/bench8.json 2> gpurun_out/r2/bench8.err; echo "c5 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/bench8.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks')}); print(d['e2e']); print(d.get('parity')); print({k:d['roofline'][k] for k in ('frac','launch_ms','dram_frac')})
PY
tail -3 gpurun_out/r2/bench8.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches8.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-facade > gpurun_out/r2/launches8.log 2>&1
tail -2 gpurun_out/r2/launches8.log | cut -c1-200
bash tools/ncu_one.sh 0 k_lattice_tma r2/tma3
