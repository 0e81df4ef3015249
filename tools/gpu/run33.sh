#!/bin/bash
# round 2, GPU call 33 (8 GPUs): the ensemble split with the temperature program handed over in one call (vx_step_ambient), slab tests on 2/4 devices of one process
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_slabbed.py -m gpu -q -k "per_device" > gpurun_out/r2/pytest33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest33.log
tail -4 gpurun_out/r2/pytest33.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --config c4 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2/bench33_c4_n$n.json 2> gpurun_out/r2/bench33_c4_n$n.err; echo "c4 n$n rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2/bench33_c4_n$n.json').read().strip().splitlines()[-1])
    print('c4 n$n', {k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['clocks'].get('reasons'))
except Exception as e: print('parse', e)
PY
done
timeout 300 python tools/slabbed_bench.py 256 256 2048 8 200 > gpurun_out/r2/slabbed33_n8.json 2> gpurun_out/r2/slabbed33_n8.err; echo "slabbed n8 rc=$?"; cat gpurun_out/r2/slabbed33_n8.json; tail -2 gpurun_out/r2/slabbed33_n8.err
