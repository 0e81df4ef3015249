"""Same-box A/B of kernel-variant builds (tools/build_variant.sh): every lib*.so named on the command line steps the same 256^3
cantilever; ms per step from CUDA events (vx_step_profile), best of 3 x 32 steps after 48 warm-up steps, twice in alternation;
the final positions of all variants must have the same bits.

    python tools/variant_sweep.py [edge] name [name ...]        names under voxelyze_b200/lib/variants/lib<name>.so
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from voxelyze_b200 import capi, scenarios

args = sys.argv[1:]
n = int(args.pop(0)) if args and args[0].isdigit() else 256
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "voxelyze_b200", "lib", "variants")
sc = scenarios.cantilever(n, n, n)
ref = None
for rep in range(2):
    for name in args:
        lib = capi.VxLib(os.path.join(root, "lib%s.so" % name))
        sim = scenarios.build(lib, sc)
        dt = sim.recommended_dt()
        sim.step(dt, 48)
        best = min(sim.step_profile(dt, 32)[0]["step"] / 32 for _ in range(3))
        tip = sim.download("pos", sim.n_voxels - 4096, 4096)
        if ref is None:
            ref = tip
        print("%-8s rep %d  %.4f ms/step  bits %s" % (name, rep, best, "same" if np.array_equal(tip, ref) else "DIFFERENT"), flush=True)
        sim.close()
