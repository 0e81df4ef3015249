#!/bin/bash
# One `ncu --set full` capture of the fused kernel on the C4 workload (4096 actuated robots).  usage: tools/ncu_c4.sh <out-name> [robots]
set -e
OUT=${1:-c4cap}; N=${2:-4096}
mkdir -p gpurun_out/$(dirname $OUT)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_lattice_tma -s 300 -c 1 -f -o gpurun_out/$OUT \
  python -c "
import sys; sys.path.insert(0,'.')
from voxelyze_b200 import capi, scenarios
lib = capi.load_product()
sim = scenarios.build(lib, scenarios.robot_ensemble($N, 10))
dt = sim.recommended_dt()
for _ in range(320):
    sim.set_temperature_all(scenarios.robot_temperature(sim.time()))
    sim.step(dt, 1)
" > gpurun_out/$OUT.log 2>&1
tail -3 gpurun_out/$OUT.log
