#!/bin/bash
# One `ncu --set full` capture of a lattice kernel variant.  usage: tools/ncu_one.sh <path> <kernel-regex> <out-name> [edge]
set -e
P=${1:-0}; K=${2:-k_lattice_tile}; OUT=${3:-cap}; N=${4:-256}
mkdir -p gpurun_out/$(dirname $OUT)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 1 -f -o gpurun_out/$OUT \
  python -c "
import sys; sys.path.insert(0,'.')
from voxelyze_b200 import capi, scenarios
lib = capi.load_product()
sim = scenarios.build(lib, scenarios.cantilever($N,$N,$N), path=$P)
dt = sim.recommended_dt()
for _ in range(30): sim.step(dt, 1)
" > gpurun_out/$OUT.log 2>&1
tail -3 gpurun_out/$OUT.log
