"""Small runs that touch every kernel of the library, meant to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py 5 7"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from voxelyze_b200 import capi, scenarios
import cases

lib = capi.load_product()
paths = [int(a) for a in sys.argv[1:] if a.isdigit()] or [0, 1, 5, 7]
ell = np.array([[i, j, k] for k in range(8) for j in range(24) for i in range(24) if j < 8 or i < 8], np.int32)
for path in paths:
    todo = [scenarios.cantilever(9, 6, 5, tip_load=20.0), scenarios.robot_ensemble(3, 5), scenarios.robot_ensemble(16, 4),     # plain, stacked and packed ensembles
            cases.BY_NAME["poisson_mixed_bilinear"].make(),                                                                        # Poisson coupling (fused: POISSON kernel)
            scenarios.plate_stack(16, 4, 2, 3, 2, tip_load=0.5),                                                                   # collisions: stale test, rebuild chain, narrowphase
            scenarios.Scenario("ell", 0.005, [capi.Material()], ell, np.zeros(len(ell), np.uint16))]                               # sparse body: brick-group list
    for sc in todo:
        sim = scenarios.build(lib, sc, path=path)
        dt = sim.recommended_dt()
        sim.step(dt, 3)
        sim.set_temperature_all(2.0)
        sim.step(dt, 18)          # one captured graph (with the conditional rebuild node when collisions are on)
        sim.download("pos"); sim.download("force_neg"); sim.state_info(8, 2); sim.download_voxel_state(0, 2)
        if sc.sim_id is None:
            sim.mesh(2, 2)
        if len(sc.ext_voxel) and (sc.ext_dof != 0).any():
            sim.linear_solve(1e-8, 400)     # static solve kernels (vx_linsolve.cuh)
        print("path", path, sc.name, "ok", sim.kernel_name()[:32], flush=True)
        sim.close()

# z-slabs of one process on the peer-store halo (GSKIP kernels, PUSH + POISSON, ghost words, vx_step_ambient, slabbed stateInfo)
if "--no-slabs" not in sys.argv:
    from test_slab_gloo import _general_scenario
    from test_slabbed import holes_scenario, poisson_scenario
    for sc in (_general_scenario(), holes_scenario(), poisson_scenario()):
        multi = scenarios.build_slabbed(lib, sc, [0, 0, 0])
        dt = multi.recommended_dt()
        for k in range(12):
            multi.set_temperature_all(0.5 * k)
            multi.step(dt, 1)
        multi.step(dt, 9)
        multi.download("pos"); multi.download("force_neg"); multi.state_info(8, 2); multi.state_info(7, 2); multi.download_voxel_state(0, 3)
        print("slabbed", sc.name, "ok halo", multi.halo_mode, multi.slab(0).kernel_name()[:40], flush=True)
        multi.close()
    sim = scenarios.build(lib, scenarios.robot_ensemble(4, 5), path=7)
    sim.step_ambient(sim.recommended_dt(), [1.0, -2.0, 3.0, 0.5, 0.0])
    print("step_ambient ok", flush=True)
    sim.close()
