#!/bin/bash
# round 2, GPU call 3: A/B of kernel variants on one box, headline parity tests, bench lines (c5 with the CPU parity leg, c4)
mkdir -p gpurun_out/r2
for v in r2a main r2a main; do
  if [ $v = main ]; then unset VX_PRODUCT_SO; else export VX_PRODUCT_SO=$PWD/voxelyze_b200/lib/variants/lib$v.so; fi
  echo "== $v" >> gpurun_out/r2/sweep3.log
  timeout 300 python tools/path_sweep.py 256 0 >> gpurun_out/r2/sweep3.log 2>&1
done
unset VX_PRODUCT_SO
cat gpurun_out/r2/sweep3.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2/pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest3.log
tail -12 gpurun_out/r2/pytest3.log
timeout 900 python bench.py --config c4 --steps 200 --warmup 20 > gpurun_out/r2/bench3_c4.json 2> gpurun_out/r2/bench3_c4.err; echo "c4 rc=$?"
cut -c1-400 gpurun_out/r2/bench3_c4.json; tail -3 gpurun_out/r2/bench3_c4.err
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench3.json 2> gpurun_out/r2/bench3.err; echo "c5 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/bench3.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','e2e')})
print(d.get('parity')); print(d.get('cpu_baseline'))
PY
tail -3 gpurun_out/r2/bench3.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2/bench3_ref.json 2> gpurun_out/r2/bench3_ref.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/r2/bench3_ref.json
