"""Models for the static solve (CVoxelyze::doLinearSolve, src/Voxelyze.cpp:243-249): shared by the CPU tests (oracle port
against the reference's own CVX_LinearSolver) and the GPU tests (device conjugate-gradient solve against the oracle).

Each entry: name -> (scenario factory, dynamic steps to run BEFORE the solve).  The pre-steps matter: fixed degrees of
freedom keep their CURRENT displacement in the solve (applyBX, src/VX_LinearSolver.cpp:288-296), so a model whose
prescribed translation / rotation has been applied by a few time steps exercises the non-homogeneous boundary terms."""
import numpy as np

import cases
from voxelyze_b200 import scenarios
from voxelyze_b200.capi import Material, DOF_ALL


def _cantilever():
    return scenarios.cantilever(10, 3, 3, tip_load=0.02)


def _two_materials_partial_fixes():
    """Striped stiff/soft block; one face fully fixed, one edge held in z only, a roller (x and rotation about y free),
    forces and moments on free voxels, a force on a FIXED dof (must be ignored, applyBX :297)."""
    ijk = scenarios.box_ijk(7, 4, 3)
    mats = [Material(E=1e6, rho=1e3), Material(E=3e7, rho=2e3, nu=0.3)]
    mat = ((ijk[:, 0] // 2 + ijk[:, 2]) % 2).astype(np.uint16)
    sc = scenarios.Scenario("two_materials_partial", 0.002, mats, ijk, mat)
    fixed = np.nonzero(ijk[:, 0] == 0)[0]
    zonly = np.nonzero((ijk[:, 0] == 6) & (ijk[:, 2] == 0))[0]
    roller = np.nonzero((ijk[:, 0] == 3) & (ijk[:, 1] == 0) & (ijk[:, 2] == 2))[0]
    load = np.nonzero((ijk[:, 0] == 6) & (ijk[:, 2] == 2))[0]
    ev = np.concatenate([fixed, zonly, roller, load]).astype(np.int32)
    dof = np.concatenate([np.full(len(fixed), DOF_ALL), np.full(len(zonly), 0x04), np.full(len(roller), 0x3F & ~0x01 & ~0x10), np.zeros(len(load))]).astype(np.uint8)
    f = np.zeros((len(ev), 3), np.float32); m = np.zeros((len(ev), 3), np.float32)
    f[len(fixed):len(fixed) + len(zonly)] = [0.001, 0.0, 5.0]            # z part sits on a fixed dof
    f[-len(load):] = [0.002, -0.001, -0.004]
    m[-len(load):] = [1e-6, -2e-6, 3e-6]
    sc.ext_voxel, sc.ext_dof, sc.ext_force, sc.ext_moment = ev, dof, f, m
    return sc


def _reversed_insertion_order():
    """Voxels created in DESCENDING x, y, z: the lower voxelsList index is the POSITIVE end of every link, which flips the
    role the reference gives the two ends in the element matrix (src/VX_LinearSolver.cpp:171-173)."""
    sc = scenarios.cantilever(6, 2, 3, tip_load=0.01)
    order = np.arange(len(sc.ijk))[::-1]
    inv = np.empty_like(order); inv[order] = np.arange(len(order))
    sc.ijk = np.ascontiguousarray(sc.ijk[order]); sc.mat = np.ascontiguousarray(sc.mat[order])
    sc.ext_voxel = inv[sc.ext_voxel].astype(np.int32)
    sc.name = "reversed_order"
    return sc


def _shuffled_insertion_order():
    sc = _two_materials_partial_fixes()
    order = np.random.default_rng(5).permutation(len(sc.ijk))
    inv = np.empty_like(order); inv[order] = np.arange(len(order))
    sc.ijk = np.ascontiguousarray(sc.ijk[order]); sc.mat = np.ascontiguousarray(sc.mat[order])
    sc.ext_voxel = inv[sc.ext_voxel].astype(np.int32)
    sc.name = "shuffled_order"
    return sc


def _members():
    """Three independent members of different loads in one handle: one block-diagonal system."""
    base = scenarios.box_ijk(5, 2, 2)
    ijk = np.tile(base, (3, 1)); sid = np.repeat(np.arange(3), len(base)).astype(np.int32)
    sc = scenarios.Scenario("members", 0.005, [Material(E=2e6, rho=1e3)], ijk, np.zeros(len(ijk), np.uint16), sim_id=sid)
    fixed = np.nonzero(ijk[:, 0] == 0)[0]; load = np.nonzero(ijk[:, 0] == 4)[0]
    ev = np.concatenate([fixed, load]).astype(np.int32)
    f = np.zeros((len(ev), 3), np.float32)
    f[len(fixed):] = np.stack([0.001 * (sid[load] + 1), np.zeros(len(load)), -0.002 * sid[load]], 1)
    sc.ext_voxel, sc.ext_dof, sc.ext_force = ev, np.concatenate([np.full(len(fixed), DOF_ALL), np.zeros(len(load))]).astype(np.uint8), f
    return sc


def _l_shape():
    ijk = np.array([[i, j, k] for k in range(2) for j in range(6) for i in range(6) if j < 2 or i < 2], np.int32)
    sc = scenarios.Scenario("ell", 0.005, [Material(E=1e6, rho=1e3)], ijk, np.zeros(len(ijk), np.uint16))
    fixed = np.nonzero(ijk[:, 0] == 5)[0]; load = np.nonzero(ijk[:, 1] == 5)[0]
    return scenarios._externals(sc, fixed, load, [0.001, 0.0, -0.002])


STATIC = {
    "cantilever": (_cantilever, 0),
    "two_materials_partial": (_two_materials_partial_fixes, 0),
    "reversed_order": (_reversed_insertion_order, 0),
    "shuffled_order": (_shuffled_insertion_order, 0),
    "members": (_members, 0),
    "ell": (_l_shape, 0),
    "prescribed_after_steps": (cases.BY_NAME["mixed_six"].make, 40),     # prescribed translation + rotation already applied
    "cantilever_after_steps": (_cantilever, 25),                          # a moving state: momenta must be zeroed, poses replaced
}


def solve(lib, name, path=0, rel_tol=1e-13):
    make, pre = STATIC[name]
    sc = make()
    sim = scenarios.build(lib, sc, path=path)
    if pre:
        sim.step(sim.recommended_dt(), pre)
    info = sim.linear_solve(rel_tol, 0)
    return sc, sim, info


def displacement_scale(sim, sc):
    pos = sim.download("pos")
    return float(np.abs(pos - sc.ijk * sc.voxel_size).max())
