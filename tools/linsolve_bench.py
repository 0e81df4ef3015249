"""Static solve (vx_linear_solve) timing: iterations, ms per iteration and the HBM rate of the two iteration kernels.

    python tools/linsolve_bench.py [nx ny nz]...      default: 64x16x16 128x32x32 256x64x64

Algorithmic bytes per voxel per iteration (csrc/vx_linsolve.cuh): k_lin_step_a reads z, p (48 B each; the six neighbours'
copies come from L1/L2), 6 neighbour indices (24), material (2) and fixed mask (1), writes p and y (48 each) = 219 B;
k_lin_step_b reads x, r, p, y, 1/diag (240) and writes x, r, z (144) = 384 B.  603 B in all."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxelyze_b200 import capi, scenarios

BYTES = 603

def main():
    a = [int(x) for x in sys.argv[1:]]
    sizes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)] or [(64, 16, 16), (128, 32, 32), (256, 64, 64)]
    lib = capi.load_product()
    for nx, ny, nz in sizes:
        sc = scenarios.cantilever(nx, ny, nz, tip_load=1.0)
        sim = scenarios.build(lib, sc)
        sim.linear_solve(1e-2, 0)                                   # warm-up: allocations, first launches
        sim.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        iters, res = sim.linear_solve(1e-10, 0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n = sim.n_voxels
        pos = sim.download("pos")
        tip = float(-(pos - sc.ijk * sc.voxel_size)[sc.ijk[:, 0] == nx - 1][:, 2].mean())
        print(json.dumps({"workload": f"cantilever {nx}x{ny}x{nz} static solve", "voxels": n, "unknowns": 6 * n, "iterations": iters, "rel_residual": res,
                          "seconds": dt, "ms_per_iteration": 1e3 * dt / max(iters, 1), "gbs": BYTES * n * iters / dt / 1e9, "tip_deflection_m": tip}), flush=True)
        sim.close()

if __name__ == "__main__":
    main()
