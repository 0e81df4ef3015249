#!/bin/bash
# round 2, GPU call 17: small-model cluster kernel (k_small_steps) -- full GPU suite, smoke, C1/C2 timings with and without it
mkdir -p gpurun_out/r2
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2/pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest17.log
tail -30 gpurun_out/r2/pytest17.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
for v in small; do
  if [ $v = nosmall ]; then export VX_NO_SMALL=1; else unset VX_NO_SMALL; fi
  echo "== $v" >> gpurun_out/r2/configs17.log
  timeout 600 python tools/config_bench.py --config c1 --steps 10000 --warmup 200 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['config'], d['kernel'][:24], round(d['ms_per_step'] * 1e3, 3), 'us/step', d['updates_per_s'])" >> gpurun_out/r2/configs17.log
done
unset VX_NO_SMALL
cat gpurun_out/r2/configs17.log
