"""Time every lattice kernel variant (vx_set_path) on the same cantilever; ms per step from CUDA events.

    python tools/path_sweep.py [scenario] [size] [paths...]   e.g.  python tools/path_sweep.py 256 0 2 1
                                                              python tools/path_sweep.py robots 4096 0 1
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelyze_b200 import capi, scenarios

NAMES = {0: "auto (warp brick, TMA)", 1: "general two-kernel", 2: "block brick 8x4x4", 3: "per-voxel fused", 4: "z-march fused", 5: "warp brick, cp.async", 6: "marching warp brick", 7: "warp brick, TMA staging"}

def main():
    what = "cantilever"
    argv = sys.argv[1:]
    if argv and not argv[0].isdigit():
        what, argv = argv[0], argv[1:]                     # robots <count> | drop <edge> | cantilever <edge>
    n = int(argv[0]) if argv else 256
    paths = [int(a) for a in argv[1:]] or [0, 2, 1]
    lib = capi.load_product()
    for path in paths:
        sc = {"cantilever": lambda: scenarios.cantilever(n, n, n), "drop": lambda: scenarios.drop_block(n),
              "robots": lambda: scenarios.robot_ensemble(n, 10), "robots1": lambda: scenarios.robot_ensemble(n, 10),
              "holes": lambda: None}[what]()
        if what == "holes":                                # n^3 block with a square through-hole and a notch: 77 % of its bounding box
            ijk = scenarios.box_ijk(n, n, n)
            keep = ~((abs(ijk[:, 0] - n // 2) < n // 5) & (abs(ijk[:, 1] - n // 2) < n // 5)) & ~((ijk[:, 0] > 3 * n // 4) & (ijk[:, 2] > 3 * n // 4))
            ijk = ijk[keep]
            sc = scenarios.Scenario("holes_%d" % n, 0.005, [capi.Material(E=1e6, rho=1e3)], ijk, __import__("numpy").zeros(len(ijk), "uint16"))
        if what == "robots1":
            sc.mat[:] = 0                                  # same geometry, one material: isolates the cost of the table look-ups
            sc.materials = sc.materials[:1]
        sim = scenarios.build(lib, sc, path=path)
        dt = sim.recommended_dt()
        sim.step(dt, 48)
        best = min(sim.step_profile(dt, 32)[0]["step"] / 32 for _ in range(3))
        print("%s %d  path %d  %-24s %.3f ms/step  %.3e voxel-updates/s" % (what, n, path, NAMES.get(path, "?"), best, sim.n_voxels / best * 1e3), flush=True)
        sim.close()

if __name__ == "__main__":
    main()
