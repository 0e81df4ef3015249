"""Time every lattice kernel variant (vx_set_path) on the same cantilever; ms per step from CUDA events.

    python tools/path_sweep.py [edge] [paths...]        e.g.  python tools/path_sweep.py 256 0 5 1
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelyze_b200 import capi, scenarios

NAMES = {0: "auto (warp brick 4x4x2)", 1: "general two-kernel", 2: "block brick 8x4x4", 3: "per-voxel fused", 4: "z-march fused", 5: "warp brick 4x4x2"}

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    paths = [int(a) for a in sys.argv[2:]] or [0, 2, 1]
    lib = capi.load_product()
    for path in paths:
        sim = scenarios.build(lib, scenarios.cantilever(n, n, n), path=path)
        dt = sim.recommended_dt()
        sim.step(dt, 48)
        best = min(sim.step_profile(dt, 32)[0]["step"] / 32 for _ in range(3))
        print("edge %d  path %d  %-24s %.3f ms/step  %.3e voxel-updates/s" % (n, path, NAMES.get(path, "?"), best, n ** 3 / best * 1e3), flush=True)
        sim.close()

if __name__ == "__main__":
    main()
