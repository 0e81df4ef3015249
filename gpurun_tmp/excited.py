import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from voxelyze_b200 import capi, scenarios
lib = capi.load_product()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for path in (0, 1):
    sc = scenarios.cantilever(n, n, n, tip_load=1.0)
    sim = scenarios.build(lib, sc, path=path)
    dt = sim.recommended_dt()
    sim.step(dt, 40)
    ms0, _ = sim.step_profile(dt, 10)
    rng = np.random.default_rng(1)
    pos = sim.download("pos"); pos += rng.standard_normal(pos.shape) * 1e-7
    sim.upload("pos", pos)
    q = sim.download("orient"); q[:, 1:] += rng.standard_normal((len(q), 3)) * 1e-5
    sim.upload("orient", q)
    sim.step(dt, 20)
    ms1, _ = sim.step_profile(dt, 10)
    print("path", path, "at rest: %.3f ms/step   excited: %.3f ms/step" % (ms0["step"] / 10, ms1["step"] / 10), flush=True)
    sim.close()
