#!/bin/bash
# round 2, GPU call 19: k_small_steps launch shapes on C1 and on a 12x12x12 block (1728 voxels)
mkdir -p gpurun_out/r2
for cfg in "0 0"; do
  set -- $cfg; export VX_SMALL_TPB=$1 VX_SMALL_CTAS=$2
  echo "== tpb $1 ctas<= $2" >> gpurun_out/r2/small19.log
  timeout 300 python tools/config_bench.py --config c1 --steps 10000 --warmup 200 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['config'], d['kernel'][:24], round(d['ms_per_step'] * 1e3, 3), 'us/step')
    elif 'rror' in l: print(l.rstrip())" >> gpurun_out/r2/small19.log
  timeout 300 python - >> gpurun_out/r2/small19.log 2>&1 <<'PY'
import sys, time; sys.path.insert(0, '.')
import torch
from voxelyze_b200 import capi, scenarios
lib = capi.load_product()
for path in (0, 7):
    sim = scenarios.build(lib, scenarios.cantilever(12, 12, 12, tip_load=1.0), path=path); dt = sim.recommended_dt()
    sim.step(dt, 200); torch.cuda.synchronize(); t0 = time.perf_counter(); sim.step(dt, 4000); torch.cuda.synchronize()
    print("12^3 path", path, sim.kernel_name()[:16], round((time.perf_counter() - t0) / 4000 * 1e6, 3), "us/step")
PY
done
cat gpurun_out/r2/small19.log
