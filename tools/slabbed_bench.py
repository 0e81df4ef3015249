"""One process, several GPUs: a cantilever lattice through vx_slabbed_* (the entry a caller of the C++ class API reaches N
GPUs through).  Host wall clock around the blocking call -- this IS the end-to-end number of that entry.

    python tools/slabbed_bench.py NX NY NZ N_DEVICES [steps]      e.g.  python tools/slabbed_bench.py 256 256 512 2
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelyze_b200 import capi, scenarios


def main():
    nx, ny, nz, nd = (int(a) for a in sys.argv[1:5])
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 200
    lib = capi.load_product()
    t0 = time.time()
    sc = scenarios.cantilever(nx, ny, nz)
    multi = scenarios.build_slabbed(lib, sc, list(range(nd)))
    build_s = time.time() - t0
    dt = multi.recommended_dt()
    multi.step(dt, 40)                                    # warm-up: graphs, tensor maps, peer mappings
    best = None
    for _ in range(3):
        t0 = time.perf_counter(); div = multi.step(dt, steps); t1 = time.perf_counter()
        assert div is None
        ms = (t1 - t0) * 1e3 / steps
        best = ms if best is None else min(best, ms)
    updates = multi.n_voxels + multi.n_links
    tip = multi.download("pos", multi.n_voxels - 1, 1)
    print(json.dumps({"tool": "slabbed_bench", "lattice": [nx, ny, nz], "devices": nd, "slabs": multi.n_slabs, "halo_mode": multi.halo_mode,
                      "steps_per_call": steps, "ms_per_step": round(best, 4), "updates_per_s": updates / best * 1e3, "build_s": round(build_s, 1),
                      "kernel": multi.slab(0).kernel_name(), "tip_z": float(tip[0, 2])}), flush=True)
    multi.close()


if __name__ == "__main__":
    main()
