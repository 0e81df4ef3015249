"""vx_slabbed_*: one model, several slabs, ONE process (include/voxelyze_b200.h).  The composition is host code shared by every
library that implements the header, so its partition logic -- slab ranges, ghost planes, externals, link numbering, gather
and scatter in the caller's numbering -- is checked here on the CPU with the restatement library and host halo copies,
bit for bit against the unsplit run of the same calls; tests/test_gpu_slabbed.py runs the same checks on the CUDA library
with the in-kernel peer stores."""
import numpy as np
import pytest

import parity
from test_slab_gloo import _general_scenario
from voxelyze_b200 import capi, scenarios

VOXEL_FIELDS = ("pos", "orient", "linmom", "angmom", "temp", "voxflags")
LINK_FIELDS = ("force_neg", "force_pos", "moment_neg", "moment_pos", "pos2", "angle1v", "angle2v", "strain", "maxstrain", "strainoffset", "stress", "linkflags")


def holes_scenario():
    """A 6 x 5 x 12 block with a window through it and a notch, created in shuffled order, hanging from its x = 0 face."""
    ijk = scenarios.box_ijk(6, 5, 12, origin=(-3, 2, -4))
    keep = ~((ijk[:, 0] >= -1) & (ijk[:, 0] <= 0) & (ijk[:, 2] >= 0) & (ijk[:, 2] <= 3)) & ~((ijk[:, 0] == 2) & (ijk[:, 1] == 6) & (ijk[:, 2] > 5))
    ijk = ijk[keep]
    ijk = ijk[np.random.default_rng(11).permutation(len(ijk))]
    mats = [capi.Material(E=1e6, rho=1e3, zeta_global=0.01), capi.Material(E=2e6, rho=1.5e3, model=capi.MODEL_BILINEAR, plastic_modulus=2e5, yield_stress=2e3, fail_stress=-1.0)]
    mat = (ijk[:, 1] & 1).astype(np.uint16)
    sc = scenarios.Scenario("slabbed_holes", 0.005, mats, ijk, mat, gravity=1.0)
    fixed = np.nonzero(ijk[:, 0] == -3)[0]
    load = np.nonzero(ijk[:, 0] == 2)[0]
    sc.ext_voxel = np.concatenate([fixed, load]).astype(np.int32)
    sc.ext_dof = np.concatenate([np.full(len(fixed), capi.DOF_ALL), np.zeros(len(load))]).astype(np.uint8)
    f = np.zeros((len(sc.ext_voxel), 3), np.float32); f[len(fixed):] = [0.0, 0.01, -0.05]
    sc.ext_force = f
    return sc


def poisson_scenario():
    """A 5 x 4 x 10 bar along z, nu = 0.3 bilinear skin around a nu = 0 core (tVoxelyze.h:806-842 turned upright), clamped at the
    bottom, its top plane pulled up by a prescribed displacement: the lateral contraction crosses every cut."""
    ijk = scenarios.box_ijk(5, 4, 10, origin=(0, 0, -2))
    ijk = ijk[np.random.default_rng(2).permutation(len(ijk))]
    core = (ijk[:, 0] >= 1) & (ijk[:, 0] <= 3) & (ijk[:, 1] >= 1) & (ijk[:, 1] <= 2)
    mats = [capi.Material(E=1e6, rho=1e3, nu=0.3, zeta_internal=1.0, zeta_global=0.2, model=capi.MODEL_BILINEAR, plastic_modulus=2e5, yield_stress=4e4),
            capi.Material(E=2e6, rho=1e3, nu=0.0, zeta_internal=1.0, zeta_global=0.2)]
    sc = scenarios.Scenario("slabbed_poisson", 0.005, mats, ijk, core.astype(np.uint16))
    bottom, top = np.nonzero(ijk[:, 2] == -2)[0], np.nonzero(ijk[:, 2] == 7)[0]
    sc.ext_voxel = np.concatenate([bottom, top]).astype(np.int32)
    sc.ext_dof = np.concatenate([np.full(len(bottom), capi.DOF_ALL), np.full(len(top), 0x04)]).astype(np.uint8)
    t = np.zeros((len(sc.ext_voxel), 3)); t[len(bottom):, 2] = 0.002
    sc.ext_translation = t
    return sc


def check_slabbed_against_whole(lib, sc, devices, steps, temperature_program=True, expect_halo=None, chunk=1, path=0):
    whole = scenarios.build(lib, sc, path=path)
    multi = scenarios.build_slabbed(lib, sc, devices)
    assert multi.n_slabs == min(len(devices), (int(sc.ijk[:, 2].max()) - int(sc.ijk[:, 2].min()) + 1) // 2)
    if expect_halo is not None:
        assert multi.halo_mode == expect_halo, multi.halo_mode
    assert multi.n_voxels == whole.n_voxels and multi.n_links == whole.n_links
    for a, b in zip(multi.links(), whole.links()):          # the reference's link creation order, restated on the host
        assert np.array_equal(a, b)
    dt = whole.recommended_dt()
    assert np.float32(dt) == np.float32(multi.recommended_dt())
    k = 0
    while k < steps:
        if temperature_program:
            t = 3.0 * np.sin(k / 10.0)
            whole.set_temperature_all(t); multi.set_temperature_all(t)
        assert whole.step(dt, chunk) is None and multi.step(dt, chunk) is None
        k += chunk
    assert whole.time() == multi.time()
    for f in VOXEL_FIELDS:
        assert parity.bit_equal(multi.download(f), whole.download(f)), f
    for f in LINK_FIELDS:
        assert parity.bit_equal(multi.download(f), whole.download(f)), f
    for info in range(10):                                   # CVoxelyze::stateInfo of the whole model (enum values of include/Voxelyze.h:48-67)
        count = whole.n_links if info in (5, 6, 7) else whole.n_voxels
        lo, hi = whole.state_info(info, 0), whole.state_info(info, 1)
        assert multi.state_info(info, 0) == lo and multi.state_info(info, 1) == hi, info
        # totals: float accumulation in list order (the reference, the restatement) against double (the CUDA library, the
        # slabbed handle) differ by the rounding of a float sum of `count` terms
        slack = 2e-7 * count * max(abs(lo), abs(hi)) + 1e-30
        a, b = multi.state_info(info, 2), whole.state_info(info, 2)
        assert abs(a - b) <= slack + 2e-6 * abs(b), (info, "total", a, b)
        a, b = multi.state_info(info, 3), whole.state_info(info, 3)
        assert abs(a - b) <= slack / count + 2e-6 * abs(b), (info, "average", a, b)
    n, nl = whole.n_voxels, whole.n_links
    for i in (0, n // 3, n - 1):                             # single elements come from the owning slab
        assert parity.bit_equal(multi.download("pos", i, 1), whole.download("pos", i, 1))
        a, b = multi.download_voxel_state(i, 1), whole.download_voxel_state(i, 1)
        assert a.tobytes() == b.tobytes()
    assert multi.download_voxel_state(0, n).tobytes() == whole.download_voxel_state(0, n).tobytes()
    if lib.backend != "reference":                         # packed link records (the reference shim has none)
        assert multi.download_link_state().tobytes() == whole.download_link_state().tobytes()
        for i in (0, nl // 2, nl - 1):
            assert multi.download_link_state(i, 1).tobytes() == whole.download_link_state(i, 1).tobytes()
    return whole, multi, dt


def test_slabbed_handle_equals_the_unsplit_run_bitwise(built):
    lib = capi.load_oracle()
    check_slabbed_against_whole(lib, _general_scenario(), [0, 0, 0], 90, expect_halo=1)


def test_slabbed_body_with_holes_and_plastic_links(built):
    lib = capi.load_oracle()
    whole, multi, dt = check_slabbed_against_whole(lib, holes_scenario(), [0] * 4, 150, temperature_program=False, expect_halo=1, chunk=10)
    assert (whole.download("linkflags") & capi.LF_YIELDED).any()


def check_state_edits(lib, devices):
    sc = _general_scenario()
    whole, multi, dt = check_slabbed_against_whole(lib, sc, devices, 40)
    n = whole.n_voxels
    rng = np.random.default_rng(3)
    pos = whole.download("pos") + 1e-6 * rng.standard_normal((n, 3))
    temp = rng.uniform(-2, 2, n).astype(np.float32)
    records = lib.backend != "reference"
    if records:
        link = whole.download_link_state(); link["strain_offset"] += np.float32(1e-5)
    for s in (whole, multi):
        s.upload("pos", pos)                                  # everything at once ...
        s.upload("linmom", np.zeros((5, 3)), first=n // 2)   # ... a short range (per element: owner + ghost copies) ...
        s.set_temperature(temp)
        if records:
            s.upload_link_state(link)
            s.upload_link_state(link[7:9], first=7)
        assert s.step(dt, 25) is None
    for f in VOXEL_FIELDS + LINK_FIELDS:
        assert parity.bit_equal(multi.download(f), whole.download(f)), f
    for s in (whole, multi):
        s.set_clock(0.25, dt)                                 # what a model that moves between handles takes along
        assert s.time() == 0.25 and s.step(dt, 3) is None
        s.reset()
        assert s.time() == 0.0
        assert s.step(dt, 10) is None
    for f in VOXEL_FIELDS:
        assert parity.bit_equal(multi.download(f), whole.download(f)), f


def test_slabbed_poisson_materials(built):
    """Poisson coupling across the cuts: a ghost copy cannot compute its Poisson strain (it lacks links), the halo brings its
    owner's.  Also with the stable time step re-evaluated every step (dt < 0), and with Poisson's ratio switched on mid-run."""
    lib = capi.load_oracle()
    sc = poisson_scenario()
    whole, multi, dt = check_slabbed_against_whole(lib, sc, [0, 0, 0], 150, temperature_program=False, expect_halo=1, chunk=5)
    assert np.abs(whole.download("pstrain")).max() > 1e-4
    assert parity.bit_equal(multi.download("pstrain"), whole.download("pstrain"))
    assert whole.step(-1.0, 7) is None and multi.step(-1.0, 7) is None
    assert whole.time() == multi.time()
    for f in VOXEL_FIELDS + LINK_FIELDS:
        assert parity.bit_equal(multi.download(f), whole.download(f)), f
    # switched on mid-run (src/VX_Link.cpp:160-166)
    sc0 = poisson_scenario(); sc0.materials[0].nu = 0.0
    whole, multi, dt = check_slabbed_against_whole(lib, sc0, [0, 0, 0], 60, temperature_program=False, chunk=5)
    for s in (whole, multi):
        s.set_materials(sc.materials)
        assert s.step(dt, 60) is None
    for f in VOXEL_FIELDS + LINK_FIELDS:
        assert parity.bit_equal(multi.download(f), whole.download(f)), f
    stiffer = poisson_scenario().materials; stiffer[0].nu = 0.4; stiffer[1].E = 3e6      # and changed again, still non-zero
    for s in (whole, multi):
        s.set_materials(stiffer)
        assert s.step(dt, 40) is None
    for f in VOXEL_FIELDS + LINK_FIELDS:
        assert parity.bit_equal(multi.download(f), whole.download(f)), f


def test_slabbed_state_edits_reach_every_copy(built):
    """vx_slabbed_upload / upload_link_state / set_temperature / reset: the ghost copies across the cuts follow, so the run that
    continues from edited state stays bit-identical to the unsplit run given the same edits."""
    check_state_edits(capi.load_oracle(), [0, 0, 0])


def test_slabbed_divergence_and_refusals(built):
    lib = capi.load_oracle()
    sc = scenarios.cantilever(6, 3, 8, tip_load=1.0)
    whole, multi = scenarios.build(lib, sc), scenarios.build_slabbed(lib, sc, [0, 0])
    dt = 40.0 * whole.recommended_dt()                        # far beyond the stable step: the run blows up
    a, b = whole.step(dt, 400), multi.step(dt, 400)
    assert a is not None and a == b
    # thin bodies run on one slab
    thin = scenarios.cantilever(6, 3, 3)
    assert scenarios.build_slabbed(lib, thin, [0, 0, 0]).n_slabs == 1
    with pytest.raises(ValueError):
        scenarios.build_slabbed(lib, scenarios.plate_stack(8, 8, 2, thick=2, gap=1), [0, 0])       # self-collisions


def test_slabbed_model_cut_again_with_another_slab_count(built):
    lib = capi.load_oracle()
    thin, tall = scenarios.cantilever(8, 4, 5, tip_load=5.0), scenarios.cantilever(8, 4, 14, tip_load=5.0)
    m = scenarios.build_slabbed(lib, thin, [0, 0, 0])
    assert m.n_slabs == 2 and m.step(m.recommended_dt(), 25) is None
    m.set_voxels(tall.ijk, tall.mat)
    m.set_externals(tall.ext_voxel, tall.ext_dof, tall.ext_force)
    assert m.n_slabs == 3
    whole = scenarios.build(lib, tall); dt = whole.recommended_dt()
    assert m.step(dt, 60) is None and whole.step(dt, 60) is None
    for f in VOXEL_FIELDS:
        assert parity.bit_equal(m.download(f), whole.download(f)), f


def test_slabbed_edge_cases(built):
    """Host logic of the slabbed handle: empty models, duplicates, partial ranges, bodies with an empty plane at a cut."""
    lib = capi.load_oracle()
    m = lib.create_slabbed(0.005, [0, 0, 0])
    m.set_materials([capi.Material(E=1e6, rho=1e3)])
    m.set_voxels(np.zeros((0, 3), np.int32), np.zeros(0, np.uint16))
    assert m.n_voxels == 0 and m.n_links == 0 and m.n_slabs == 0 and m.step(1e-5, 3) is None and m.time() == 0.0
    with pytest.raises(capi.VxError) as e:
        m.set_voxels(np.array([[0, 0, 0], [0, 0, 1], [0, 0, 0]], np.int32), np.zeros(3, np.uint16))
    assert e.value.code == -5
    sc = scenarios.cantilever(5, 4, 9, tip_load=2.0)
    with pytest.raises(capi.VxError):
        scenarios.build_slabbed(lib, sc, [0, 0]).set_externals([sc.n_voxels], [capi.DOF_ALL])
    # partial ranges in the caller's numbering, across the cuts
    whole, multi = scenarios.build(lib, sc), scenarios.build_slabbed(lib, sc, [0, 0, 0])
    dt = whole.recommended_dt()
    assert whole.step(dt, 30) is None and multi.step(dt, 30) is None
    n, nl = whole.n_voxels, whole.n_links
    for first, count in ((3, 5), (n // 2 - 20, 47), (n - 30, 30), (0, n)):
        assert parity.bit_equal(multi.download("pos", first, count), whole.download("pos", first, count))
        assert multi.download_voxel_state(first, count).tobytes() == whole.download_voxel_state(first, count).tobytes()
    for first, count in ((1, 6), (nl // 3, 100), (nl - 9, 9)):
        assert parity.bit_equal(multi.download("strain", first, count), whole.download("strain", first, count))
        assert parity.bit_equal(multi.download("force_pos", first, count), whole.download("force_pos", first, count))
    kick = 1e-7 * np.random.default_rng(1).standard_normal((60, 3))
    for s in (whole, multi):                                   # a range of 60 voxels: more than the per-element path takes
        s.upload("linmom", kick, first=n // 2 - 30)
        assert s.step(dt, 20) is None
    for f in VOXEL_FIELDS:
        assert parity.bit_equal(multi.download(f), whole.download(f)), f
    # two bodies above each other with an empty plane where the cut falls: the ghost planes there are empty
    ijk = np.concatenate([scenarios.box_ijk(4, 3, 4), scenarios.box_ijk(4, 3, 4, origin=(0, 0, 5))])
    two = scenarios.Scenario("two_bodies", 0.005, [capi.Material(E=1e6, rho=1e3)], ijk, np.zeros(len(ijk), np.uint16), gravity=1.0)
    two.ext_voxel = np.nonzero(ijk[:, 0] == 0)[0].astype(np.int32); two.ext_dof = np.full(len(two.ext_voxel), capi.DOF_ALL, np.uint8)
    check_slabbed_against_whole(lib, two, [0, 0, 0], 40, temperature_program=False)


@pytest.mark.parametrize("which", ["general", "poisson"])
def test_state_moves_between_handles_with_its_clock(built, which):
    """The complete persistent state of a run: voxel fields, link records, the clock (time, CVX_Voxel::previousDt) and -- with
    Poisson materials -- the cached Poisson strains (a fully fixed voxel never refreshes its cache, src/VX_Voxel.cpp:167-172).
    Moved from one handle into three slabs and back into one handle, the run goes on with the bits of the run that never moved."""
    lib = capi.load_oracle()
    sc = _general_scenario() if which == "general" else poisson_scenario()
    a = scenarios.build(lib, sc); dt = a.recommended_dt()
    assert a.step(dt, 80) is None
    src = a
    for dst in (scenarios.build_slabbed(lib, sc, [0, 0, 0]), scenarios.build(lib, sc)):
        for f in VOXEL_FIELDS:
            dst.upload(f, src.download(f))
        dst.upload_link_state(src.download_link_state())
        if which == "poisson":
            dst.upload("pstrain", src.download("pstrain"))
        dst.set_clock(src.time(), dt)
        assert a.step(dt, 40) is None and dst.step(dt, 40) is None
        for f in VOXEL_FIELDS + LINK_FIELDS:
            assert parity.bit_equal(dst.download(f), a.download(f)), (f, type(dst).__name__)
        src = dst
