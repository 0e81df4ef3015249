"""Static solve, CPU side: the oracle port's restatement of CVX_LinearSolver against the reference's own
CVX_LinearSolver::solve (calculateA / applyBX / postResults unmodified, PARDISO's two entry points supplied by a banded
Cholesky in oracle/ref_shim.cpp), and against the closed-form answers the reference's tests quote for a single bond."""
import numpy as np
import pytest

import static_cases
from voxelyze_b200 import scenarios
from voxelyze_b200.capi import Material, DOF_ALL, VxError


@pytest.mark.parametrize("name", list(static_cases.STATIC))
def test_oracle_solve_matches_the_references_solver(oracle, reference, name):
    sc, o, _ = static_cases.solve(oracle, name)
    _, r, _ = static_cases.solve(reference, name)
    scale = static_cases.displacement_scale(r, sc)
    assert scale > 0
    # two direct factorisations of the same matrix: agreement to rounding (conditioning ~1e6 at these sizes)
    assert np.abs(o.download("pos") - r.download("pos")).max() <= 1e-9 * scale
    assert np.abs(o.download("orient") - r.download("orient")).max() <= 1e-9
    for f in ("linmom", "angmom"):
        assert not o.download(f).any() and not r.download(f).any()


def _bond(lib, force=(0, 0, 0), moment=(0, 0, 0)):
    """tVoxelyze.h:142-198 singleBondFixedFree: two 1 mm voxels, E = 1 MPa, the first fixed."""
    sc = scenarios.Scenario("bond", 0.001, [Material(E=1e6, rho=1e3)], np.array([[0, 0, 0], [1, 0, 0]], np.int32), np.zeros(2, np.uint16))
    sc.ext_voxel = np.array([0, 1], np.int32); sc.ext_dof = np.array([DOF_ALL, 0], np.uint8)
    sc.ext_force = np.array([[0, 0, 0], force], np.float32); sc.ext_moment = np.array([[0, 0, 0], moment], np.float32)
    sim = scenarios.build(lib, sc)
    sim.linear_solve()
    pos = sim.download("pos")[1] - [0.001, 0, 0]
    q = sim.download("orient")[1]
    return pos, 2 * q[1:]                                  # small angles: rotation vector = 2 * vector part


@pytest.mark.parametrize("which", ["oracle", "reference"])
def test_single_bond_closed_forms(request, which):
    """The steady states the reference's dynamic tests converge to (tVoxelyze.h:155-197): 1e-6 m axial under 1e-3 N,
    6e-3 rad and 4e-6 m under a transverse 1e-3 N, 1.2e-5 rad under 1e-9 N m bending, 3e-6 rad... (torsion: a2 = E L^3 / 12)."""
    lib = request.getfixturevalue(which)
    u, th = _bond(lib, force=(1e-3, 0, 0))
    assert abs(u[0] - 1e-6) < 1e-12 and abs(u[1]) < 1e-15 and abs(u[2]) < 1e-15
    u, th = _bond(lib, force=(0, 1e-3, 0))
    assert abs(u[1] - 4e-6) < 1e-11 and abs(th[2] - 6e-3) < 1e-8
    u, th = _bond(lib, force=(0, 0, 1e-3))
    assert abs(u[2] - 4e-6) < 1e-11 and abs(th[1] + 6e-3) < 1e-8
    u, th = _bond(lib, moment=(0, 0, 1e-9))
    assert abs(th[2] - 1.2e-5) < 1e-11 and abs(u[1] - 6e-9) < 1e-14
    u, th = _bond(lib, moment=(1e-9, 0, 0))
    assert abs(th[0] - 1.2e-5) < 1e-11                    # a2 = E L^3 / 12 = 8.33e-5 N m / rad


@pytest.mark.parametrize("which", ["oracle", "reference"])
def test_unheld_model_is_reported(request, which):
    """No fixed voxel: the matrix is singular, the solve fails (PARDISO error -4) and the state is untouched."""
    lib = request.getfixturevalue(which)
    sc = scenarios.cantilever(4, 2, 2, tip_load=0.01)
    sc.ext_dof[:] = 0
    sim = scenarios.build(lib, sc)
    before = sim.download("pos")
    with pytest.raises(VxError):
        sim.linear_solve()
    assert np.array_equal(before, sim.download("pos"))
