"""Catalogue of small parity cases shared by the golden generator, the oracle-vs-reference
tests (CPU) and the CUDA parity tests (GPU).  Most restate a reference gtest
(test/tVoxelyze.h of the reference, line cited per case) as a flat scenario."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from voxelyze_b200 import scenarios
from voxelyze_b200.capi import Material, DOF_ALL, MODEL_BILINEAR, MODEL_DATA
from voxelyze_b200.scenarios import Scenario


@dataclass
class Case:
    name: str
    make: Callable[[], Scenario]
    steps: int
    program: Optional[Callable] = None     # program(sim, k, t) before step k
    smooth: bool = True                    # smooth nu=0 case: 1e-9 tolerance applies on GPU
    tol: float = 1e-9                      # GPU-vs-oracle relative tolerance (position metric)
    exact_flags: bool = True


def _mk(name, ijk, mats, mat=None, size=0.001, **kw) -> Scenario:
    ijk = np.asarray(ijk, np.int32).reshape(-1, 3)
    m = np.zeros(len(ijk), np.uint16) if mat is None else np.asarray(mat, np.uint16)
    return Scenario(name, size, mats, ijk, m, **kw)


def _ext(sc, vox, dof, force=None, moment=None, tr=None, rot=None):
    sc.ext_voxel = np.asarray(vox, np.int32)
    sc.ext_dof = np.asarray(dof, np.uint8)
    n = len(sc.ext_voxel)
    sc.ext_force = None if force is None else np.asarray(force, np.float32).reshape(n, 3)
    sc.ext_moment = None if moment is None else np.asarray(moment, np.float32).reshape(n, 3)
    sc.ext_translation = None if tr is None else np.asarray(tr, np.float64).reshape(n, 3)
    sc.ext_rotation = None if rot is None else np.asarray(rot, np.float64).reshape(n, 3)
    return sc


def _box(nx, ny, nz):
    return [[i, j, k] for i in range(nx) for j in range(ny) for k in range(nz)]


STD = dict(E=1e6, rho=1e3)


def single_bond(axis_dir=(1, 0, 0), force=(1e-3, 0, 0), moment=(0, 0, 0), dof=0x3E):
    """tVoxelyze.h:63-113 test2Vox: fixed voxel + one loaded voxel."""
    def make():
        sc = _mk("single_bond", [[0, 0, 0], list(axis_dir)], [Material(zeta_internal=1.0, zeta_global=0.2, **STD)])
        return _ext(sc, [0, 1], [DOF_ALL, dof], force=[[0, 0, 0], list(force)], moment=[[0, 0, 0], list(moment)])
    return make


def combined_damping():
    """tVoxelyze.h:336-379: 4x3x3 block, x=0 fixed, x=3 loaded."""
    ijk = _box(4, 3, 3)
    sc = _mk("combined_damping", ijk, [Material(zeta_internal=1.0, zeta_global=0.05, **STD)])
    a = np.array(ijk)
    fx, ld = np.nonzero(a[:, 0] == 0)[0], np.nonzero(a[:, 0] == 3)[0]
    f = np.zeros((len(fx) + len(ld), 3), np.float32); f[len(fx):, 2] = 1e-6
    return _ext(sc, np.concatenate([fx, ld]), [DOF_ALL] * len(fx) + [0] * len(ld), force=f)


def large_deformation():
    """tVoxelyze.h:445-473: big load flips the link into large-angle mode."""
    sc = _mk("large_deformation", [[0, 0, 0], [1, 0, 0]], [Material(zeta_internal=1.0, zeta_global=0.2, **STD)])
    return _ext(sc, [0, 1], [DOF_ALL, 0], force=[[0, 0, 0], [-0.2, 0, 0.2]])


def mixed_six():
    """all three link axes, large rotations, partial DOF fixes, prescribed translation+rotation."""
    ijk = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 1, 1]]
    sc = _mk("mixed_six", ijk, [Material(zeta_internal=1.0, zeta_global=0.2, **STD)])
    return _ext(sc, [0, 5, 4], [DOF_ALL, 0x08 | 0x02, 0],
                force=[[0, 0, 0], [0.05, -0.03, 0.08], [0.01, 0.02, -0.05]],
                moment=[[0, 0, 0], [1e-5, 2e-5, -1e-5], [0, 0, 3e-5]],
                tr=[[0, 0, 0], [0, 1e-4, 0], [0, 0, 0]], rot=[[0.1, 0.2, -0.1], [0.3, 0, 0], [0, 0, 0]])


def multi_material():
    """tVoxelyze.h:648-684 multiSimple2: alternating soft/stiff pairs along x."""
    mats = [Material(E=1e6, rho=1e3, zeta_global=0.01, zeta_internal=1.0), Material(E=1e9, rho=1e3, zeta_global=0.01, zeta_internal=1.0)]
    ijk = [[i, 0, 0] for i in range(8)]
    sc = _mk("multi_material", ijk, mats, mat=[(i // 2) % 2 for i in range(8)])
    return _ext(sc, [0, 7], [DOF_ALL, 0], force=[[0, 0, 0], [1e-3, 0, 0]])


def bilinear_yield():
    """tVoxelyze.h:845-889 deformableMaterial: bilinear 1e6/5e5/1e5, loaded past yield."""
    ijk = _box(5, 3, 3)
    sc = _mk("bilinear_yield", ijk, [Material(model=MODEL_BILINEAR, E=1e6, plastic_modulus=5e5, yield_stress=1e5, rho=1e3,
                                               zeta_internal=1.0, zeta_global=0.2)])
    a = np.array(ijk)
    fx, ld = np.nonzero(a[:, 0] == 0)[0], np.nonzero(a[:, 0] == 4)[0]
    f = np.zeros((len(fx) + len(ld), 3), np.float32); f[len(fx):, 0] = 0.2
    return _ext(sc, np.concatenate([fx, ld]), [DOF_ALL] * len(fx) + [0] * len(ld), force=f)


def _bilinear_program(sim, k, t):
    if k == 400:     # erase the forces (tVoxelyze.h:873-878)
        a = np.array(_box(5, 3, 3))
        fx, ld = np.nonzero(a[:, 0] == 0)[0], np.nonzero(a[:, 0] == 4)[0]
        sim.set_externals(np.concatenate([fx, ld]), [DOF_ALL] * len(fx) + [0] * len(ld))


def data_curve_fail():
    """data-curve material pulled until links fail; second material linear with failure stress."""
    md = Material(model=MODEL_DATA, strain=[0.01, 0.02, 0.04, 0.08], stress=[1e4, 1.8e4, 3e4, 4e4], rho=1e3, zeta_global=0.1)
    ml = Material(E=1e6, rho=1e3, fail_stress=3.5e4, zeta_global=0.1)
    ijk = [[i, j, 0] for i in range(6) for j in range(2)]
    a = np.array(ijk)
    sc = _mk("data_curve_fail", ijk, [md, ml], mat=(a[:, 0] % 2))
    fx, ld = np.nonzero(a[:, 0] == 0)[0], np.nonzero(a[:, 0] == 5)[0]
    f = np.zeros((len(fx) + len(ld), 3), np.float32); f[len(fx):, 0] = 0.045
    return _ext(sc, np.concatenate([fx, ld]), [DOF_ALL] * len(fx) + [0] * len(ld), force=f)


def temperature_bimorph():
    """tVoxelyze.h:989-1032: CTE bimorph bends when heated."""
    m1 = Material(E=1e6, rho=1e3, zeta_internal=1.0, zeta_global=0.15, cte=0.01)
    m2 = Material(E=1e7, rho=1e3, zeta_internal=1.0, zeta_global=0.15)
    ijk, mat = [], []
    for i in range(3):
        ijk += [[i, 0, 0], [i, 0, 1]]; mat += [0, 1]
    sc = _mk("temperature_bimorph", ijk, [m1, m2], mat=mat, temperature=5.0)
    return _ext(sc, [0, 1], [DOF_ALL, DOF_ALL])


def friction_slide():
    """tVoxelyze.h:1034-1167: single voxel on the floor, pushed past static friction, then released."""
    sc = _mk("friction_slide", [[0, 0, 0]], [Material(E=1e6, rho=1e3, mu_static=1.0, mu_kinetic=0.1, zeta_global=1.0)],
             gravity=1.0, floor=True)
    return sc


def _friction_program(sim, k, t):
    nf = np.float32(1e3 * 1e-9 * 9.80665)
    if k == 50:
        sim.set_externals([0], [0], force=[[1.1 * nf, 0, 0]])
    if k == 80:
        sim.set_externals([0], [0], force=[[0, 0, 0]])


def poisson_block():
    """tVoxelyze.h:715-769 poissonsLarge: 9x2x2, nu=0.3, prescribed end displacement."""
    ijk = _box(9, 2, 2)
    a = np.array(ijk)
    sc = _mk("poisson_block", ijk, [Material(E=1e6, rho=1e3, nu=0.3, zeta_internal=1.0, zeta_global=0.2)])
    fx, mv = np.nonzero(a[:, 0] == 0)[0], np.nonzero(a[:, 0] == 8)[0]
    tr = np.zeros((len(fx) + len(mv), 3)); tr[len(fx):, 0] = np.float32(1e-3)
    return _ext(sc, np.concatenate([fx, mv]), [DOF_ALL] * (len(fx) + len(mv)), tr=tr)


def poisson_mixed_bilinear():
    """tVoxelyze.h:806-944 flavour: nu=0.3 bilinear + nu=0 linear, force loaded."""
    mb = Material(model=MODEL_BILINEAR, E=1e6, plastic_modulus=5e5, yield_stress=1e5, rho=1e3, nu=0.3, zeta_internal=1.0, zeta_global=0.2)
    m0 = Material(E=1e6, rho=1e3, nu=0.0, zeta_internal=1.0, zeta_global=0.3)
    ijk = _box(5, 3, 3)
    a = np.array(ijk)
    sc = _mk("poisson_mixed_bilinear", ijk, [mb, m0], mat=(a[:, 0] >= 3).astype(int))
    fx, ld = np.nonzero(a[:, 0] == 0)[0], np.nonzero(a[:, 0] == 4)[0]
    f = np.zeros((len(fx) + len(ld), 3), np.float32); f[len(fx):, 0] = 0.2
    return _ext(sc, np.concatenate([fx, ld]), [DOF_ALL] * len(fx) + [0] * len(ld), force=f)


def collide_two():
    """tVoxelyze.h:1169-1197: voxel dropped on a fixed voxel, self collision holds it up."""
    sc = _mk("collide_two", [[0, 0, 0], [0, 0, 2]], [Material(E=1e6, rho=1e6)], gravity=1.0, floor=True, collisions=True)
    return _ext(sc, [0], [DOF_ALL])


def robots_program(sim, k, t):
    sim.set_temperature_all(scenarios.robot_temperature(t))


# the scenarios of the z-slab tests (tests/test_slab_gloo.py, tests/test_slabbed.py), as whole models: what the slabbed runs are
# compared with is itself pinned to the reference
def _slab_general():
    from test_slab_gloo import _general_scenario
    return _general_scenario()


def _slab_holes():
    from test_slabbed import holes_scenario
    return holes_scenario()


def _slab_poisson():
    from test_slabbed import poisson_scenario
    return poisson_scenario()


def _slab_program(sim, k, t):
    sim.set_temperature_all(3.0 * np.sin(k / 10.0))


CASES = [
    Case("c1_cantilever", scenarios.cantilever, 10000),
    Case("single_bond_axial", single_bond(), 300),
    Case("single_bond_y_moment", single_bond((0, 1, 0), (0, 0, 0), (1e-9, 0, 0), 0x37), 300),
    Case("single_bond_z_shear", single_bond((0, 0, -1), (1e-3, 0, 0), (0, 0, 0), 0x3E), 300),
    Case("combined_damping", combined_damping, 1000),
    Case("large_deformation", large_deformation, 400, smooth=False, tol=1e-8),
    Case("mixed_six", mixed_six, 1500, smooth=False, tol=1e-7),
    Case("multi_material", multi_material, 3000),
    Case("bilinear_yield", bilinear_yield, 650, program=_bilinear_program),
    Case("data_curve_fail", data_curve_fail, 3000, smooth=False, tol=1e-7),
    Case("temperature_bimorph", temperature_bimorph, 500),
    Case("friction_slide", friction_slide, 300, program=_friction_program),
    Case("drop_block_6", lambda: scenarios.drop_block(6), 4000, tol=1e-7),
    Case("robots_3x4", lambda: scenarios.robot_ensemble(3, 4), 600, program=robots_program, tol=1e-7),
    Case("poisson_block", poisson_block, 300, smooth=False, tol=1e-6),
    Case("poisson_mixed_bilinear", poisson_mixed_bilinear, 500, smooth=False, tol=1e-6),
    Case("collide_two", collide_two, 300, smooth=False, tol=1e-6),
    Case("plates_16x4x2", lambda: scenarios.plate_stack(16, 4, 2, 3, 2, tip_load=0.5), 3000, smooth=False, tol=1e-6),
    Case("slab_general", _slab_general, 120, program=_slab_program, smooth=False, tol=1e-7),
    Case("slab_holes", _slab_holes, 150, smooth=False, tol=1e-7),
    Case("slab_poisson", _slab_poisson, 150, smooth=False, tol=1e-6),
]
BY_NAME = {c.name: c for c in CASES}
