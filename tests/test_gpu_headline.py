"""Parity at (or near) the sizes BASELINE.json is quoted on -- the CUDA product through the C-ABI against the
UNMODIFIED reference (oracle/_ref, OpenMP build; nu = 0 so the OpenMP build is bit-identical to the serial one,
SURVEY.md section 8c) or, where the reference build is absent, the oracle port.

  * C5 pattern at 128^3, 100 steps                      ("100 steps at C5a"; 256^3 x 20 steps is carried by bench.py's
                                                          `parity` field, which runs next to the timed region on the GPU box)
  * C3's validated scale model, 60 000 steps             (32 x 8 x (3 + 2 gap + 3): 1 636 watched pairs, 120 yielded and
                                                          48 failed links)
  * one C4 robot (10^3, three materials, CTE program), 2 000 steps
/root/reference is never read here: the checkers are the prebuilt oracle/_ref/*.so and oracle/liboracle_port.so."""
import os

import numpy as np
import pytest

import parity
from voxelyze_b200 import capi, scenarios

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cpu(built):
    """The reference's own code on all cores (OpenMP build), else its serial build, else the oracle port."""
    if os.path.exists(capi.REF_OMP_SO):
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)       # read by libgomp when the library is loaded
        return capi.load_reference(omp=True)
    if os.path.exists(capi.REF_SO):
        return capi.load_reference()
    return capi.load_oracle()


def test_c5_pattern_128_cubed_100_steps_against_the_reference(product, cpu):
    """Every voxel of a 128^3 cantilever (2 097 152 voxels, 6 242 304 links) after 100 steps: positions within 1e-9 of the
    displacement scale, quaternion components within 1e-9, link flags equal."""
    n = 128
    sc = scenarios.cantilever(n, n, n, tip_load=1.0)
    g, dt, dg = parity.run(product, sc, 100)
    c, dtc, dc = parity.run(cpu, sc, 100, dt=dt)
    assert g.active_path() == 2 and dg is None and dc is None
    assert np.float32(c.recommended_dt()) == np.float32(dt)
    fields = ["pos", "orient", "linmom", "angmom"]
    err = parity.rel_errors(parity.snapshot(g, fields), parity.snapshot(c, fields), sc)
    assert err["pos"] <= 1e-9 and err["orient"] <= 1e-9, err
    assert err["linmom"] <= 1e-6 and err["angmom"] <= 1e-6, err
    assert np.array_equal(g.download("linkflags") & 0xD, c.download("linkflags") & 0xD)
    # the wave front has travelled 100 voxels from the loaded face in 100 steps: most of the lattice took part
    moved = np.abs(g.download("pos") - sc.ijk * sc.voxel_size).max(axis=1) > 0
    assert moved.mean() > 0.3, moved.mean()


def test_c3_scale_model_60000_steps_pairs_and_flags_exact(product, built):
    """SURVEY 8d's validated scale model of C3: a cantilever plate bends onto the slab below it, yields and partly fails
    (1 636 watched pairs, 120 yielded and 48 failed links on the unmodified reference).  Watched pair set and the
    yielded / failed flag of every link: exact.  Trajectory: centre of mass and kinetic energy.  The checker is the
    reference's SERIAL build (regenerateCollisions is serial in the reference anyway)."""
    cpu_serial = capi.load_reference() if os.path.exists(capi.REF_SO) else capi.load_oracle()
    sc = scenarios.plate_stack(32, 8, 2, 3, 2, tip_load=1.0)
    steps = 60000
    g = scenarios.build(product, sc); dt = g.recommended_dt()
    c = scenarios.build(cpu_serial, sc)
    assert np.float32(c.recommended_dt()) == np.float32(dt)
    com_err = 0.0
    for k in range(0, steps, 10000):                 # compare the trajectory on the way, not only its end
        assert g.step(dt, 10000) is None and c.step(dt, 10000) is None
        pg, pc = g.download("pos"), c.download("pos")
        com_err = max(com_err, float(np.abs(pg.mean(axis=0) - pc.mean(axis=0)).max()))
    fg, fc = g.download("linkflags"), c.download("linkflags")
    pairs_g, pairs_c = g.collision_pairs(), c.collision_pairs()
    yielded, failed = int(((fc & 4) != 0).sum()), int(((fc & 8) != 0).sum())
    assert (len(pairs_c), yielded, failed) == (1636, 120, 48), (len(pairs_c), yielded, failed)      # SURVEY 8d, measured on the reference
    assert np.array_equal(pairs_g[np.lexsort(pairs_g.T[::-1])], pairs_c[np.lexsort(pairs_c.T[::-1])]), "watched pair set"
    assert np.array_equal(fg & 0xC, fc & 0xC), "yielded / failed flags"
    scale = float(np.abs(pc - sc.ijk * sc.voxel_size).max())
    assert com_err <= 1e-6 * scale, (com_err, scale)
    mass = g.voxmat(0)["mass"]
    ke_g = float((g.download("linmom") ** 2).sum()) / (2 * mass)
    ke_c = float((c.download("linmom") ** 2).sum()) / (2 * mass)
    assert abs(ke_g - ke_c) <= 1e-4 * max(ke_c, 1e-30) + 1e-18, (ke_g, ke_c)      # residual motion after 60 000 steps of contact and yielding


@pytest.mark.parametrize("path", [7, 0], ids=["fused", "auto-small"])
def test_c4_one_robot_2000_steps_against_the_reference(product, cpu, path):
    """One robot of the C4 population (10^3 voxels, materials by hash, CTE +/-0.01, floor, gravity, friction), ambient
    temperature 20 sin(2 pi 40 t) set before every step, 2 000 steps.  Smooth until the first floor contact; with Coulomb
    friction afterwards the stated bound is 1e-6 (SURVEY 8d parity metrics, C2/C4)."""
    sc = scenarios.robot_ensemble(1, 10, first_seed=7)
    program = lambda sim, k, t: sim.set_temperature_all(scenarios.robot_temperature(t))
    g, dt, dg = parity.run(product, sc, 2000, program=program, path=path)
    c, _, dc = parity.run(cpu, sc, 2000, dt=dt, program=program)
    assert dg is None and dc is None and g.active_path() == parity.layout(path, sc.n_voxels)
    sg, so = parity.snapshot(g), parity.snapshot(c)
    err = parity.rel_errors(sg, so, sc)
    assert err["pos"] <= 1e-6 and err["orient"] <= 1e-6, err
    assert np.array_equal(sg["voxflags"], so["voxflags"])            # static-friction state machine
    assert np.array_equal(sg["temp"], so["temp"])
    assert np.array_equal(sg["linkflags"] & 0xD, so["linkflags"] & 0xD)


@pytest.mark.parametrize("case_name,steps", [("data_curve_fail", 300), ("large_deformation", 400), ("multi_material", 300), ("temperature_bimorph", 300)])
def test_surface_mesh_matches_the_references_mesh_render(product, built, case_name, steps):
    """the device mesh (vertices averaged from deformed voxel corners, quad normals, all colour schemes)
    against the reference's own CVX_MeshRender driven through oracle/_ref.  Topology (vertex numbering, quads, quad owners):
    exact.  Vertices: 1e-6 of the voxel size (the corner arithmetic is float; voxel state differs by <= 1e-9).  Colours: 2e-3
    (jet map of a value divided by a float-accumulated maximum)."""
    import cases
    if not os.path.exists(capi.REF_SO):
        pytest.skip("oracle/_ref not built")
    ref = capi.load_reference()
    sc = cases.BY_NAME[case_name].make()
    rgba = [[255, 40, 0, 255], [0, 128, 255, 255], [10, 200, 30, 128], [77, 77, 77, 255]][:len(sc.materials)]
    sims = []
    for lib in (product, ref):
        s = scenarios.build(lib, sc); dt = sc.dt or s.recommended_dt()
        program = cases.BY_NAME[case_name].program
        if program is None:
            s.step(dt, steps)
        else:
            t = 0.0
            for k in range(steps):
                program(s, k, t); s.step(dt, 1); t = float(np.float32(t) + np.float32(dt))
        s.mesh_set_material_colors(rgba)
        sims.append(s)
    g, r = sims
    for coloring, state in ((0, 0), (1, 0), (2, 2), (2, 0), (2, 6), (2, 5), (2, 7), (2, 8)):
        a, b = g.mesh(coloring, state), r.mesh(coloring, state)
        assert np.array_equal(a["quads"], b["quads"]) and np.array_equal(a["quad_voxel"], b["quad_voxel"])
        assert a["vertices"].shape == b["vertices"].shape and len(a["quads"]) > 0
        assert np.abs(a["vertices"] - b["vertices"]).max() <= 1e-6 * sc.voxel_size, (coloring, state)
        assert np.abs(a["normals"] - b["normals"]).max() <= 1e-4, (coloring, state)
        assert np.abs(a["colors"] - b["colors"]).max() <= 2e-3, (coloring, state, np.abs(a["colors"] - b["colors"]).max())
