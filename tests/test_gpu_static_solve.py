"""Static solve on the device (vx_linear_solve: matrix-free preconditioned conjugate gradients, csrc/vx_linsolve.cuh) against
the oracle's direct solve of the reference's matrix (tests/test_static_solve.py pins that one to CVX_LinearSolver itself)."""
import numpy as np
import pytest

import parity
import static_cases
from voxelyze_b200 import scenarios
from voxelyze_b200.capi import VxError

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", [0, 1], ids=["auto", "general"])
@pytest.mark.parametrize("name", list(static_cases.STATIC))
def test_device_solve_matches_the_oracle(product, oracle, name, path):
    sc, g, (iters, res) = static_cases.solve(product, name, path=path)
    _, o, _ = static_cases.solve(oracle, name)
    assert iters > 0 and res <= 1e-13
    scale = static_cases.displacement_scale(o, sc)
    # iterative against direct, condition numbers up to ~1e6: 1e-8 of the largest displacement
    assert np.abs(g.download("pos") - o.download("pos")).max() <= 1e-8 * scale, name
    assert np.abs(g.download("orient") - o.download("orient")).max() <= 1e-8, name
    for f in ("linmom", "angmom"):
        assert not g.download(f).any()
    # link state is left as it was (postResults touches voxels only, src/VX_LinearSolver.cpp:336-347)
    for f in ("strain", "force_neg"):
        a, b = g.download(f), o.download(f)
        assert np.abs(a - b).max() <= 1e-6 * max(np.abs(b).max(), 1e-30), (name, f)


def test_two_solves_give_the_same_bits(product):
    a = static_cases.solve(product, "two_materials_partial")[1]
    b = static_cases.solve(product, "two_materials_partial")[1]
    for f in ("pos", "orient"):
        assert np.array_equal(a.download(f), b.download(f))


def test_mid_size_beam_against_the_direct_solve(product, oracle):
    sc = scenarios.cantilever(24, 8, 8, tip_load=0.5)
    g = scenarios.build(product, sc, path=7); o = scenarios.build(oracle, sc)
    iters, res = g.linear_solve(1e-12, 0)
    o.linear_solve()
    assert g.active_path() == 2
    scale = static_cases.displacement_scale(o, sc)
    assert np.abs(g.download("pos") - o.download("pos")).max() <= 1e-7 * scale
    assert np.abs(g.download("orient") - o.download("orient")).max() <= 1e-7


def test_solution_is_an_equilibrium_of_the_time_stepper(product):
    """Ties the solver to the hot path: under a small load the static solution is a rest state of doTimeStep -- stepping on from
    it moves nothing beyond the geometric non-linearity the linear solve leaves out (the beam shortens by ~deflection^2 / length:
    4e-5 of the deflection here)."""
    sc = scenarios.cantilever(20, 4, 4, tip_load=2e-4)
    sc.materials[0].zeta_global = 0.05
    g = scenarios.build(product, sc)
    g.linear_solve(1e-13, 0)
    p0 = g.download("pos")
    scale = static_cases.displacement_scale(g, sc)
    dt = g.recommended_dt()
    assert g.step(dt, 400) is None
    assert np.abs(g.download("pos") - p0).max() <= 1e-4 * scale
    # and the dynamic relaxation from rest converges to the same state
    h = scenarios.build(product, sc)
    h.step(dt, 40000)
    assert np.abs(h.download("pos") - p0).max() <= 5e-3 * scale      # (lightly damped: still ringing a little after 40 000 steps)


def test_large_beam_linearity_and_beam_theory(product):
    """128 x 16 x 16 voxels (32 768 voxels, 196 608 unknowns): doubling the load doubles the displacement; the tip deflection
    is within 5 % of Euler-Bernoulli + shear (F L^3 / 3EI, I = (16 h)^4 / 12)."""
    n, w, F, h, E = 128, 16, 0.8, 0.005, 1e6
    tips = []
    for load in (F, 2 * F):
        sc = scenarios.cantilever(n, w, w, tip_load=load)
        g = scenarios.build(product, sc)
        iters, res = g.linear_solve(1e-11, 0)
        assert res <= 1e-11
        pos = g.download("pos")
        tips.append(pos - sc.ijk * sc.voxel_size)
        g.close()
    assert np.abs(tips[1] - 2 * tips[0]).max() <= 1e-7 * np.abs(tips[1]).max()
    tip = -tips[0][sc.ijk[:, 0] == n - 1][:, 2].mean()
    L = (n - 1) * h; I = (w * h) ** 4 / 12.0
    assert abs(tip / (F * L ** 3 / (3 * E * I)) - 1.0) < 0.05


def test_unheld_model_fails_and_leaves_the_state(product):
    sc = scenarios.cantilever(6, 3, 3, tip_load=0.01)
    sc.ext_dof[:] = 0
    g = scenarios.build(product, sc)
    before = parity.snapshot(g)
    with pytest.raises(VxError):
        g.linear_solve(1e-10, 3000)
    after = parity.snapshot(g)
    for f in before:
        assert parity.bit_equal(before[f], after[f]), f


def test_stepping_after_a_solve_follows_the_oracle(product, oracle):
    sc, g, _ = static_cases.solve(product, "two_materials_partial")
    _, o, _ = static_cases.solve(oracle, "two_materials_partial")
    dt = g.recommended_dt()
    g.step(dt, 60); o.step(dt, 60)
    err = parity.rel_errors(parity.snapshot(g), parity.snapshot(o), sc)
    assert err["pos"] <= 1e-6 and err["orient"] <= 1e-6, err
