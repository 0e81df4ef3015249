#!/bin/bash
for so in gpurun_tmp/t_*.so; do
  VX_PRODUCT_SO=$PWD/$so python bench.py --steps 64 --warmup 20 --no-cpu-baseline > /tmp/o.txt 2>&1
  python - "$so" <<'PY'
import sys,json
so=sys.argv[1]; ok=False
for l in open('/tmp/o.txt'):
    if l.startswith('{'):
        d=json.loads(l); ok=True; print(so, 'ms/step %.3f'%d['ms_per_step'], 'kernel %.3f'%d['roofline']['kernel_ms_per_step']['step'])
if not ok: print(so, 'FAILED:', open('/tmp/o.txt').read()[-300:].replace('\n',' | '))
PY
done
