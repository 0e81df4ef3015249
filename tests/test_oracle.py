"""Pins the oracle (CPU restatement) -- CPU only.

1. against the committed golden vectors (outputs of the unmodified reference), bit for bit;
2. against oracle/_ref (the unmodified reference compiled here) when it is present, bit for bit,
   including derived material tables and the reference's own known-answer values
   (test/tVoxelyze.h of the reference)."""
import os

import numpy as np
import pytest

import cases
import parity
from voxelyze_b200 import capi, scenarios
from voxelyze_b200.capi import Material, DOF_ALL, MODEL_BILINEAR, MODEL_DATA

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", cases.CASES, ids=lambda c: c.name)
def test_oracle_matches_golden_bitwise(oracle, case):
    gold = np.load(os.path.join(GOLDEN, case.name + ".npz"))
    sc = case.make()
    sim, dt, div = parity.run(oracle, sc, case.steps, program=case.program)
    assert np.float32(dt) == gold["dt"]
    assert (-1 if div is None else div) == int(gold["diverged"])
    snap = parity.snapshot(sim)
    for f, v in snap.items():
        assert parity.bit_equal(v, gold[f]), f
    if sc.collisions:
        assert np.array_equal(sim.collision_pairs(), gold["pairs"])


@pytest.mark.parametrize("case", [c for c in cases.CASES if c.steps <= 3000], ids=lambda c: c.name)
def test_oracle_matches_live_reference_bitwise(oracle, reference, case):
    sc = case.make()
    a, dta, da = parity.run(reference, sc, case.steps, program=case.program)
    b, dtb, db = parity.run(oracle, sc, case.steps, program=case.program)
    assert dta == dtb and da == db
    sa, sb = parity.snapshot(a), parity.snapshot(b)
    for f in sa:
        assert parity.bit_equal(sa[f], sb[f]), f
    assert np.array_equal(np.stack(a.links()), np.stack(b.links()))


MATS = [
    Material(),
    Material(E=1e9, rho=2e3, nu=0.3, cte=0.02, mu_static=1.0, mu_kinetic=0.5, zeta_internal=0.5, zeta_global=0.1, zeta_collision=0.3),
    Material(model=MODEL_BILINEAR, E=1e6, plastic_modulus=5e5, yield_stress=1e5, fail_stress=2e5, rho=1e3),
    Material(model=MODEL_BILINEAR, E=1e7, plastic_modulus=2e6, yield_stress=4e4, rho=1e3, nu=0.2),
    Material(model=MODEL_DATA, strain=[0.01, 0.02, 0.04, 0.08], stress=[1e4, 1.8e4, 3e4, 4e4], rho=1e3),
    Material(E=1e6, rho=1e3, fail_stress=3.5e4, ext_scale=(1.5, 1.0, 0.5)),
]


def _tables(lib):
    s = lib.create(0.005)
    s.set_materials(MATS)
    vox = [s.voxmat(i) for i in range(len(MATS))]
    link = {(a, b): s.linkmat(a, b) for a in range(len(MATS)) for b in range(a, len(MATS))}
    s.close()
    return vox, link


def _same(x, y):
    if isinstance(x, np.ndarray):
        return parity.bit_equal(x, y)
    if isinstance(x, list):
        return x == y
    return (x == y) or (x != x and y != y)


def test_material_tables_match_reference(oracle, reference):
    """CVX_MaterialVoxel / CVX_MaterialLink derived constants (tVX_Material.h, tVX_MaterialLink.h)."""
    ov, ol = _tables(oracle)
    rv, rl = _tables(reference)
    for a, b in zip(ov, rv):
        for k in a:
            assert _same(a[k], b[k]), k
    for key in ol:
        for k in ol[key]:
            assert _same(ol[key][k], rl[key][k]), (key, k)


def test_known_answers_of_the_reference_tests(oracle):
    """A few closed-form values the reference's gtests assert (test/tVoxelyze.h)."""
    # singleBondFixedFree :144 -- 1e-3 N axial on a 1 mm, 1 MPa link -> 1e-6 m
    c = cases.BY_NAME["single_bond_axial"]
    sim, dt, _ = parity.run(oracle, c.make(), 300)
    assert abs(sim.download("pos")[1, 0] - 0.001 - 1e-6) < 1e-6 * 1e-4
    # combinedDamping :363 -- 1.742e-8 +- 1e-10
    c = cases.BY_NAME["combined_damping"]
    sc = c.make()
    sim, dt, _ = parity.run(oracle, sc, 1000)
    i = int(np.nonzero((sc.ijk == [3, 0, 0]).all(1))[0][0])
    assert abs(sim.download("pos")[i, 2] - 1.742e-8) < 1e-10
    # multiSimple2 :676 -- 3.5035e-6 +- 1e-10 after 10000 steps
    c = cases.BY_NAME["multi_material"]
    sim, dt, _ = parity.run(oracle, c.make(), 10000)
    assert abs(sim.download("pos")[7, 0] - 0.007 - 3.5035e-6) < 1e-10
    # deformableMaterial :886 -- plastic set 4e-4 +- 1e-7
    c = cases.BY_NAME["bilinear_yield"]
    sc = c.make()
    sim, dt, _ = parity.run(oracle, sc, 650, program=c.program)
    i = int(np.nonzero((sc.ijk == [4, 1, 1]).all(1))[0][0])
    assert abs(sim.download("pos")[i, 0] - 0.004 - 4e-4) < 1e-7
    # temperature :1021 -- bimorph tip 2.55e-5 +- 1e-8
    c = cases.BY_NAME["temperature_bimorph"]
    sc = c.make()
    sim, dt, _ = parity.run(oracle, sc, 500)
    i = int(np.nonzero((sc.ijk == [2, 0, 0]).all(1))[0][0])
    assert abs(sim.download("pos")[i, 2] - 2.55e-5) < 1e-8
    # largeDeformationDamping :469 -- z = 9.5587e-4 +- 1e-7 after 200 steps
    c = cases.BY_NAME["large_deformation"]
    sim, dt, _ = parity.run(oracle, c.make(), 200)
    assert abs(sim.download("pos")[1, 2] - 9.5587e-4) < 1e-7
    assert not (sim.download("linkflags")[0] & capi.LF_SMALL_ANGLE)
    # collisions :1196 -- dropped voxel rests above the fixed one
    c = cases.BY_NAME["collide_two"]
    sim, dt, _ = parity.run(oracle, c.make(), 150)
    assert sim.download("pos")[1, 2] > 0.001


def test_c1_anchor(oracle):
    """SURVEY.md section 8c anchor of config C1 (measured on the unmodified reference)."""
    sc = scenarios.cantilever()
    sim, dt, _ = parity.run(oracle, sc, 10000)
    assert np.float32(dt) == np.float32(2.5164607e-05)
    i = int(np.nonzero((sc.ijk == [19, 0, 0]).all(1))[0][0])
    pos, q = sim.download("pos")[i], sim.download("orient")[i]
    assert pos[0] == 0.088760750681525377 and pos[2] == -0.02348399744544306
    assert pos[1] == -5.5293048677981086e-12
    assert q[0] == 0.98278022892661421


def test_material_errors_mirror_reference(oracle):
    s = oracle.create(0.001)
    with pytest.raises(capi.VxError) as e:
        s.set_materials([Material(E=-1.0)])
    assert "Young's modulus must be positive" in str(e.value)
    with pytest.raises(capi.VxError) as e:
        s.set_materials([Material(model=MODEL_BILINEAR, E=1e6, plastic_modulus=2e6, yield_stress=1e5)])
    assert "Plastic modulus" in str(e.value)


def test_edge_cases(oracle):
    s = oracle.create(0.001)
    s.set_materials([Material()])
    s.set_voxels(np.zeros((0, 3), np.int32), np.zeros(0, np.uint16))     # empty lattice
    assert s.n_voxels == 0 and s.n_links == 0 and s.recommended_dt() == 0.0
    assert s.step(1e-5, 3) is None
    s.set_voxels([[5, 5, 5]], [0])                                          # single voxel, no links
    assert s.n_links == 0 and s.recommended_dt() > 0
    with pytest.raises(capi.VxError):
        s.set_voxels([[0, 0, 0], [0, 0, 0]], [0, 0])                        # duplicate
    with pytest.raises(capi.VxError):
        s.set_voxels([[40000, 0, 0]], [0])                                  # index does not fit a short
    s.set_voxels([[0, 0, 0], [1, 0, 0], [-1, 0, 0]], [0, 0, 0])            # negative indices, creation order links
    vn, vp, ax = s.links()
    assert list(vn) == [0, 2] and list(vp) == [1, 0] and list(ax) == [0, 0]
    assert s.step(0.0, 5) is None and s.time() == 0.0                       # dt == 0 is a no-op


def test_oracle_follows_the_reference_when_poissons_ratio_is_switched_on_mid_run(oracle, reference):
    """Material edits between steps (SURVEY 8b hazard 2): nu 0 -> 0.3 after 150 steps, then 150 more; bit for bit."""
    import copy

    def run(lib):
        sc = scenarios.cantilever(8, 3, 3, tip_load=20.0)
        s = scenarios.build(lib, sc); dt = s.recommended_dt()
        s.step(dt, 150)
        m = copy.copy(sc.materials[0]); m.nu = 0.3
        s.set_materials([m])
        dt2 = s.recommended_dt()
        s.step(dt2 * 0.5, 150)
        return parity.snapshot(s), dt2
    (a, dta), (b, dtb) = run(oracle), run(reference)
    assert dta == dtb
    for f in a:
        assert parity.bit_equal(a[f], b[f]), f


@pytest.mark.parametrize("which", ["slab_general", "slab_poisson", "bilinear_yield", "large_deformation"])
def test_a_run_of_the_unmodified_reference_moves_between_objects(reference, oracle, which):
    """What the C-ABI's state-carrying calls (vx_upload, vx_upload_link_state, vx_set_clock; VX_F_PSTRAIN with Poisson materials)
    rest on: these records ARE the persistent state of the reference's objects.  A run of the unmodified reference is read out
    through them, written into a freshly built CVoxelyze, and both go on bit for bit -- and so does the restatement, fed with
    the reference's records."""
    case = cases.BY_NAME[which]
    sc = case.make()
    a = scenarios.build(reference, sc); dt = a.recommended_dt()
    first = case.steps // 2
    assert a.step(dt, first) is None
    state = {f: a.download(f) for f in parity.VOXEL_FIELDS}
    links, pstrain, clock = a.download_link_state(), a.download("pstrain"), a.time()
    movers = [scenarios.build(reference, sc), scenarios.build(oracle, sc)]
    for b in movers:
        for f, v in state.items():
            b.upload(f, v)
        b.upload_link_state(links)
        b.upload("pstrain", pstrain)
        b.set_clock(clock, dt)
    assert a.step(dt, case.steps - first) is None
    sa = parity.snapshot(a)
    for b in movers:
        assert b.step(dt, case.steps - first) is None and b.time() == a.time()
        sb = parity.snapshot(b)
        for f in sa:
            assert parity.bit_equal(sa[f], sb[f]), (f, b.L.backend)
