#!/bin/bash
# round 2, GPU call 5 (2 GPUs): slab tests, the reference's gtests on the GPU, N=2 bench with the whole-lattice bitwise check
mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/smi5.txt
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_dropin_cpp.py -m gpu -q --durations=5 > gpurun_out/r2/pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest5.log
tail -15 gpurun_out/r2/pytest5.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench5_n2.json 2> gpurun_out/r2/bench5_n2.err; echo "n2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2/bench5_n2.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','slab_bitwise','slab_check')}); print(d['roofline'].get('kernel_ms_per_rank'), d['clocks'])
except Exception as e: print("parse", e)
PY
tail -5 gpurun_out/r2/bench5_n2.err
