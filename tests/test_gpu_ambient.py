"""vx_step_ambient: an ambient temperature program in one call == setAmbientTemperature + doTimeStep in turn (BASELINE config C4's
thermally actuated robots), bit for bit -- state, the link forces a download recomputes from the last step's inputs, and the
divergence semantics."""
import numpy as np
import pytest

import parity
from voxelyze_b200 import capi, scenarios, slab

pytestmark = pytest.mark.gpu


def _program(sim, dt, n):
    t, temps = np.float32(sim.time()), []
    for _ in range(n):
        temps.append(scenarios.robot_temperature(float(t)))
        t = np.float32(t + np.float32(dt))
    return temps


@pytest.mark.parametrize("path", [0, 7, 1], ids=["auto", "fused-tma", "general"])
def test_ambient_program_equals_per_step_calls_bitwise(product, path):
    sc = scenarios.robot_ensemble(6, 6)
    a, b = scenarios.build(product, sc, path=path), scenarios.build(product, sc, path=path)
    dt = a.recommended_dt()
    for chunk in (1, 37, 100, 2):
        temps = _program(a, dt, chunk)
        for t in temps:
            a.set_temperature_all(t)
            assert a.step(dt, 1) is None
        assert b.step_ambient(dt, temps) is None
        assert a.time() == b.time()
        sa, sb = parity.snapshot(a), parity.snapshot(b)
        for f in sa:
            assert parity.bit_equal(sa[f], sb[f]), (f, chunk)
    assert np.abs(a.download("temp")).max() > 1.0


def test_ensemble_runner_steps_the_same_either_way(product):
    """bench.py --config c4: the timed leg hands the program over in one call, the e2e leg makes the per-step calls."""
    sc = scenarios.robot_ensemble(8, 10)
    a, b = slab.EnsembleRunner(scenarios.build(product, sc)), slab.EnsembleRunner(scenarios.build(product, sc))
    dt = a.recommended_dt()
    for _ in range(60):
        assert a.step(dt, 1) is None
    assert b.step(dt, 60) is None
    for f in ("pos", "orient", "linmom", "angmom", "temp", "voxflags"):
        assert parity.bit_equal(a.sim.download(f), b.sim.download(f)), f


def test_ambient_program_divergence(product):
    sc = scenarios.cantilever(12, 4, 6, tip_load=1.0)
    sc.materials[0].cte = 0.01
    a, b = scenarios.build(product, sc, path=7), scenarios.build(product, sc, path=7)
    dt = 40.0 * a.recommended_dt()
    temps = [float(k % 5) for k in range(300)]
    div = None
    for k, t in enumerate(temps):
        a.set_temperature_all(t)
        if a.step(dt, 1) is not None:
            div = k
            break
    assert div is not None and b.step_ambient(dt, temps) == div
    sa, sb = parity.snapshot(a, parity.VOXEL_FIELDS), parity.snapshot(b, parity.VOXEL_FIELDS)
    for f in sa:
        assert parity.bit_equal(sa[f], sb[f]), f
