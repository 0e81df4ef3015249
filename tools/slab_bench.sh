#!/bin/bash
# usage: tools/slab_bench.sh <n_gpus> [steps]   -- 512^3 z-slab bench through bench.py under torchrun
N=${1:-2}; STEPS=${2:-40}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps $STEPS --warmup 5 --no-cpu-baseline "${@:3}"
