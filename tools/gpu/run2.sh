#!/bin/bash
# round 2, GPU call 2: L2-prefetch distance sweep, parity tests, ncu capture
mkdir -p gpurun_out/r2
for pf in 0 148 296 444 592; do
  echo "== VX_PF_AHEAD=$pf" >> gpurun_out/r2/sweep2.log
  VX_PF_AHEAD=$pf timeout 300 python tools/path_sweep.py 256 0 >> gpurun_out/r2/sweep2.log 2>&1
done
cat gpurun_out/r2/sweep2.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest2.log
tail -4 gpurun_out/r2/pytest2.log
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2/bench2.json 2> gpurun_out/r2/bench2.err
python -c "import json; d=json.load(open('gpurun_out/r2/bench2.json')); print(d['ms_per_step'], d['roofline']['launch_ms'], d['clocks'])"
bash tools/ncu_one.sh 0 k_lattice_tma r2/tma2
