"""Small lattice runs on every fused kernel variant, meant to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py 5 7"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelyze_b200 import capi, scenarios

lib = capi.load_product()
paths = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4, 5, 6, 7]
for path in paths:
    for sc in (scenarios.cantilever(9, 6, 5, tip_load=20.0), scenarios.robot_ensemble(3, 5)):
        sim = scenarios.build(lib, sc, path=path)
        dt = sim.recommended_dt()
        sim.step(dt, 3)
        sim.step(dt, 18)          # one captured graph
        sim.download("pos"); sim.download("force_neg"); sim.state_info(8, 2)
        print("path", path, sc.name, "ok", sim.kernel_name()[:32], flush=True)
        sim.close()
