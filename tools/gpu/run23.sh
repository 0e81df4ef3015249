#!/bin/bash
# round 2, GPU call 23 (8 GPUs): scaling lines of both decompositions (C5b z-slabs, C4 ensemble) at 8/4/2(/1) GPUs
mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/smi23.txt
run() { # n, outfile, args...
  local n=$1 out=$2; shift 2
  if [ $n = 1 ]; then timeout 1200 python bench.py --gpus 1 "$@" > gpurun_out/r2/$out.json 2> gpurun_out/r2/$out.err
  else timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@" > gpurun_out/r2/$out.json 2> gpurun_out/r2/$out.err; fi
  echo "$out rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2/$out.json').read().strip().splitlines()[-1])
    print('$out', {k:d.get(k) for k in ('value','ms_per_step','slab_bitwise')}, d['roofline'].get('kernel_ms_per_rank'), d['clocks'].get('reasons'), d['clocks'].get('samples_in_timed_region'))
except Exception as e: print('$out parse', e)
PY
}
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q > gpurun_out/r2/pytest23.log 2>&1; echo "slab pytest rc=$?"; tail -3 gpurun_out/r2/pytest23.log
run 8 bench23_c5_n8_driver --steps 20 --warmup 5
run 8 bench23_c5_n8 --steps 200 --warmup 20 --no-cpu-baseline
run 4 bench23_c5_n4 --steps 200 --warmup 20 --no-cpu-baseline
run 2 bench23_c5_n2 --steps 200 --warmup 20 --no-cpu-baseline
run 8 bench23_c4_n8 --config c4 --steps 200 --warmup 20 --no-cpu-baseline
run 4 bench23_c4_n4 --config c4 --steps 200 --warmup 20 --no-cpu-baseline
run 2 bench23_c4_n2 --config c4 --steps 200 --warmup 20 --no-cpu-baseline
run 1 bench23_c4_n1 --config c4 --steps 200 --warmup 20 --no-cpu-baseline
run 1 bench23_c5_n1 --steps 200 --warmup 20 --no-cpu-baseline
tail -3 gpurun_out/r2/bench23_c5_n8_driver.err
