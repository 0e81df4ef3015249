"""Parity tests proper: the CUDA product (through the C-ABI) against the oracle on the same
seeded inputs.  Run on the GPU box with `pytest -m gpu`.

Tolerances (BASELINE.json north_star): positions <= 1e-9 relative to the displacement scale
and quaternion components <= 1e-9 on smooth nu = 0 cases; cases that go through sin/cos/acos/
pow (large-angle links, Poisson) or contact carry the looser per-case tolerance of
tests/cases.py.  Link flags (small-angle, yielded, failed) and divergence steps must match
exactly.  /root/reference is never read here."""
import os

import numpy as np
import pytest

import cases
import parity
from voxelyze_b200 import capi, scenarios
from voxelyze_b200.capi import Material

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

NOT_BUILT = set()     # cases needing features that are not on the GPU yet


@pytest.mark.parametrize("path", [0, 1, 7], ids=["auto", "general", "fused"])
@pytest.mark.parametrize("case", cases.CASES, ids=lambda c: c.name)
def test_cuda_matches_oracle(product, oracle, case, path):
    """auto: these models are small, so the small-model cluster kernel (all steps of a call in one launch) unless
    self-collisions are on; general: the one-step kernels of the general layout; fused: the lattice kernel wherever the
    bounding box is at most 8x the voxel count."""
    if case.name in NOT_BUILT:
        pytest.skip("feature not on the GPU yet")
    sc = case.make()
    g, dtg, dg = parity.run(product, sc, case.steps, program=case.program, path=path)
    o, dto, do = parity.run(oracle, sc, case.steps, program=case.program)
    assert g.launch_count() > 0
    assert dtg == dto, "recommended time step"
    assert dg == do, "divergence step"
    sg, so = parity.snapshot(g), parity.snapshot(o)
    err = parity.rel_errors(sg, so, sc)
    assert err["pos"] <= case.tol, err
    assert err["orient"] <= case.tol, err
    for f in ("pos2", "angle2v", "force_neg", "moment_neg", "strain", "stress"):
        if f in err:
            assert err[f] <= max(case.tol * 1e3, 1e-6), (f, err)
    # momenta decay to rounding noise at equilibrium: compare on the natural momentum scale
    # m * (displacement scale / dt) instead of their own (vanishing) magnitude
    mass = max(g.voxmat(i)["mass"] for i in range(len(sc.materials)))
    nominal = sc.ijk.astype(np.float64) * sc.voxel_size
    p_scale = mass * max(float(np.max(np.abs(so["pos"] - nominal))), 1e-300) / dtg
    assert float(np.max(np.abs(sg["linmom"] - so["linmom"]))) <= max(case.tol * 1e3, 1e-6) * p_scale, err
    assert np.array_equal(sg["linkflags"] & 0xD, so["linkflags"] & 0xD), "small-angle / yielded / failed flags"
    assert np.array_equal(sg["voxflags"], so["voxflags"])
    assert np.array_equal(sg["temp"], so["temp"])
    if sc.collisions:
        pg, po = g.collision_pairs(), o.collision_pairs()
        assert np.array_equal(pg[np.lexsort(pg.T[::-1])], po[np.lexsort(po.T[::-1])]), "collision pair set"


@pytest.mark.parametrize("name", ["c1_cantilever", "bilinear_yield", "temperature_bimorph"])
def test_cuda_matches_golden_reference_output(product, name):
    """Directly against the committed outputs of the unmodified reference."""
    case = cases.BY_NAME[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = case.make()
    g, dt, _ = parity.run(product, sc, case.steps, program=case.program)
    assert np.float32(dt) == gold["dt"]
    err = parity.rel_errors(parity.snapshot(g), {k: gold[k] for k in gold.files}, sc)
    assert err["pos"] <= 1e-9 and err["orient"] <= 1e-9, err


def test_material_tables_bitwise(product, oracle):
    from test_oracle import MATS, _tables, _same
    pv, pl = _tables(product)
    ov, ol = _tables(oracle)
    for a, b in zip(pv, ov):
        for k in a:
            assert _same(a[k], b[k]), k
    for key in pl:
        for k in pl[key]:
            assert _same(pl[key][k], ol[key][k]), (key, k)


def test_layout_selection(product):
    """auto: models of at most 700 voxels without self-collisions are stepped by the small-model cluster kernel on the general
    layout (one launch per vx_step call); everything else runs fused wherever its bounding box is at most 8x its voxel
    count.  vx_set_path forces either."""
    small = scenarios.build(product, scenarios.cantilever(6, 3, 3))
    assert small.active_path() == 1 and "k_small_steps" in small.kernel_name()
    assert scenarios.build(product, scenarios.cantilever(6, 3, 3), path=1).active_path() == 1
    assert "k_link" in scenarios.build(product, scenarios.cantilever(6, 3, 3), path=1).kernel_name()
    assert scenarios.build(product, scenarios.cantilever(6, 3, 3), path=7).active_path() == 2
    big = scenarios.build(product, scenarios.cantilever(40, 8, 8))                 # 2560 voxels
    assert big.active_path() == 2 and "k_small_steps" in scenarios.build(product, scenarios.cantilever(40, 8, 8), path=3).kernel_name()
    assert scenarios.build(product, scenarios.robot_ensemble(2, 3), path=7).active_path() == 2
    assert scenarios.build(product, scenarios.robot_ensemble(40, 4)).active_path() == 2           # 2560 voxels in 40 members
    assert scenarios.build(product, cases.BY_NAME["temperature_bimorph"].make(), path=7).active_path() == 2
    # a box with holes (6 of 2x2x2 cells) is filled up with inert cells and runs fused; a sparse shape does not
    six = scenarios.build(product, cases.BY_NAME["mixed_six"].make(), path=7)
    assert six.active_path() == 2 and six.n_voxels == 6
    assert scenarios.build(product, cases.BY_NAME["mixed_six"].make(), path=1).active_path() == 1
    ell = [[i, 0, 0] for i in range(5)] + [[0, j, 0] for j in range(1, 5)]
    sparse = scenarios.Scenario("ell", 0.001, [Material()], np.array(ell, np.int32), np.zeros(len(ell), np.uint16))
    assert scenarios.build(product, sparse, path=7).active_path() == 2   # 9 voxels in 25 cells: a single body may be as sparse as 1 in 8
    diag = [[i, i, i] for i in range(6)]
    very_sparse = scenarios.Scenario("diag", 0.001, [Material()], np.array(diag, np.int32), np.zeros(len(diag), np.uint16))
    assert scenarios.build(product, very_sparse, path=7).active_path() == 1      # 6 voxels in 216 cells: general layout
    assert scenarios.build(product, cases.BY_NAME["poisson_block"].make(), path=7).active_path() == 2  # nu != 0: fused too (k_lattice_tma<.., POISSON>)
    assert scenarios.build(product, cases.BY_NAME["poisson_mixed_bilinear"].make(), path=7).active_path() == 2
    # self-collisions switched on: a small model moves to the fused layout (its captured step graphs carry the collision kernels)
    col = scenarios.build(product, scenarios.plate_stack(16, 4, 2, 3, 2, tip_load=0.5))
    assert col.active_path() == 2


def test_fused_and_general_paths_agree_bitwise(product):
    """Same physics functions, same summation order: the two device layouts give identical bits."""
    sc = scenarios.cantilever(12, 5, 4, tip_load=30.0)
    snaps = {}
    for path in (0, 1, 3, 5, 7):   # auto (= 3 at this size), general one-step kernels, small-model cluster kernel, warp bricks staged by cp.async / by TMA
        sim, dt, _ = parity.run(product, sc, 700, path=path)
        snaps[path] = parity.snapshot(sim)
    for path in (1, 3, 5, 7):
        for f in snaps[0]:
            assert parity.bit_equal(snaps[0][f], snaps[path][f]), (path, f)


@pytest.mark.parametrize("path", [0, 5, 7], ids=["auto", "warpbrick-cpasync", "warpbrick-tma"])
def test_fused_kernels_odd_sizes(product, oracle, path):
    """Lattice edges that are not multiples of the 4x4x2 bricks or of their 2x2x2 groups: partial bricks,
    padded groups, zero-filled TMA boxes."""
    sc = scenarios.cantilever(33, 6, 35, tip_load=200.0)
    g, dt, _ = parity.run(product, sc, 40, path=path)
    o, _, _ = parity.run(oracle, sc, 40, dt=dt)
    assert g.active_path() == 2
    err = parity.rel_errors(parity.snapshot(g), parity.snapshot(o), sc)
    assert err["pos"] <= 1e-9 and err["orient"] <= 1e-9, err


def test_diverging_step_semantics(product, oracle):
    """A step whose links exceed strain 100 returns VX_DIVERGED, advances links but not voxels
    (src/Voxelyze.cpp:263-269), on both device layouts and inside the one-launch call of the small-model kernel."""
    c = cases.BY_NAME["data_curve_fail"]
    sc = c.make()
    for path in (0, 1, 7):
        g, dt, dg = parity.run(product, sc, 2900, path=path)
        o, _, do = parity.run(oracle, sc, 2900)
        assert dg == do and dg is not None
        sg, so = parity.snapshot(g), parity.snapshot(o)
        err = parity.rel_errors(sg, so, sc)
        assert err["pos"] <= 1e-7 and err["strain"] <= 1e-6, (path, err)
        assert abs(g.time() - o.time()) <= 1e-9


@pytest.mark.parametrize("path", [0, 1], ids=["fused", "general"])
def test_diverging_step_with_collisions_enabled(product, oracle, path):
    """with self-collisions on, the step that diverges must not touch the watch lists or the contact forces (the
    reference returns before updateCollisions, src/Voxelyze.cpp:265-271) -- also inside a multi-step call on the fused path,
    where the divergence flag of a step lives in the flag of its generation."""
    sc = cases.BY_NAME["data_curve_fail"].make()
    sc.collisions = True
    g, dt, dg = parity.run(product, sc, 2900, path=path)
    o, _, do = parity.run(oracle, sc, 2900)
    assert dg == do and dg is not None
    assert g.active_path() == (1 if path == 1 else 2)
    sg, so = parity.snapshot(g), parity.snapshot(o)
    err = parity.rel_errors(sg, so, sc)
    assert err["pos"] <= 1e-7 and err["strain"] <= 1e-6, err
    pg, po = g.collision_pairs(), o.collision_pairs()
    assert np.array_equal(pg[np.lexsort(pg.T[::-1])] if len(pg) else pg, po[np.lexsort(po.T[::-1])] if len(po) else po)
    # a second call after the divergence: still diverged at once, nothing moves
    before = g.download("pos")
    assert g.step(dt, 5) == o.step(dt, 5) == 0
    assert np.array_equal(before, g.download("pos"))


@pytest.mark.parametrize("path", [0, 1, 7], ids=["auto", "general", "fused"])
def test_determinism_two_runs_bit_equal(product, path):
    sc = scenarios.cantilever(16, 6, 5)
    a, _, _ = parity.run(product, sc, 500, path=path)
    b, _, _ = parity.run(product, sc, 500, path=path)
    sa, sb = parity.snapshot(a), parity.snapshot(b)
    for f in sa:
        assert np.array_equal(sa[f], sb[f]), f


@pytest.mark.parametrize("path", [7, 1, 0], ids=["fused", "general", "auto-small"])
def test_graph_and_single_steps_agree(product, path):
    """vx_step(n) uses captured CUDA graphs (fused, general) or one launch for the whole call (small-model kernel); one-step
    calls do not.  Same bits either way."""
    sc = scenarios.cantilever(10, 3, 3)
    a = scenarios.build(product, sc, path=path); dt = a.recommended_dt()
    a.step(dt, 100)
    b = scenarios.build(product, sc, path=path)
    for _ in range(100):
        b.step(dt, 1)
    sa, sb = parity.snapshot(a), parity.snapshot(b)
    for f in sa:
        assert np.array_equal(sa[f], sb[f]), f
    assert a.time() == b.time()


def test_creation_order_does_not_change_results(product, oracle):
    """Voxels handed over in a shuffled order: same physics, results returned in caller order."""
    sc = scenarios.cantilever(8, 3, 3)
    rng = np.random.default_rng(7)
    perm = rng.permutation(sc.n_voxels)
    inv = np.argsort(perm)
    sh = scenarios.Scenario("shuffled", sc.voxel_size, sc.materials, sc.ijk[perm], sc.mat[perm],
                            ext_voxel=inv[sc.ext_voxel].astype(np.int32), ext_dof=sc.ext_dof, ext_force=sc.ext_force)
    g, dt, _ = parity.run(product, sh, 300)
    o, _, _ = parity.run(oracle, sh, 300)
    err = parity.rel_errors(parity.snapshot(g), parity.snapshot(o), sh)
    assert err["pos"] <= 1e-9 and err["orient"] <= 1e-9, err
    assert np.array_equal(np.stack(g.links()), np.stack(o.links()))


def test_edge_cases_on_device(product):
    s = product.create(0.001)
    s.set_materials([Material()])
    s.set_voxels(np.zeros((0, 3), np.int32), np.zeros(0, np.uint16))
    assert s.n_voxels == 0 and s.recommended_dt() == 0.0 and s.step(1e-5, 3) is None
    s.set_voxels([[5, 5, 5]], [0])
    assert s.n_links == 0 and s.recommended_dt() > 0
    assert s.step(1e-6, 20) is None
    with pytest.raises(capi.VxError):
        s.set_voxels([[0, 0, 0], [0, 0, 0]], [0, 0])
    s.set_voxels([[0, 0, 0], [1, 0, 0], [-1, 0, 0]], [0, 0, 0])
    vn, vp, ax = s.links()
    assert list(vn) == [0, 2] and list(vp) == [1, 0]
    t = s.time()                           # a voxel set replaced mid-run: simulation time goes on (src/Voxelyze.cpp:422-498)
    assert t == pytest.approx(20e-6, rel=1e-4)
    assert s.step(0.0, 5) is None and s.time() == t          # dt == 0: no step (src/Voxelyze.cpp:253)
    s.reset()
    assert s.time() == 0.0


def test_reset_restores_initial_state(product, oracle):
    sc = scenarios.cantilever(8, 2, 2)
    g = scenarios.build(product, sc); o = scenarios.build(oracle, sc)
    dt = g.recommended_dt()
    for s in (g, o):
        s.step(dt, 200); s.reset(); s.step(dt, 100)
    err = parity.rel_errors(parity.snapshot(g), parity.snapshot(o), sc)
    assert err["pos"] <= 1e-9 and err["orient"] <= 1e-9, err
    assert abs(g.time() - o.time()) < 1e-12


def test_upload_download_roundtrip(product):
    sc = scenarios.cantilever(6, 3, 2)
    s = scenarios.build(product, sc)
    rng = np.random.default_rng(3)
    for f, c in (("pos", 3), ("orient", 4), ("linmom", 3), ("angmom", 3)):
        data = rng.standard_normal((sc.n_voxels, c))
        s.upload(f, data)
        assert np.array_equal(s.download(f), data)
        assert np.array_equal(s.download(f, 5, 7), data[5:12])
    t = rng.standard_normal(sc.n_voxels).astype(np.float32)
    s.set_temperature(t)
    assert np.array_equal(s.download("temp"), t)


def test_mid_size_lattice_few_steps(product, oracle):
    """40^3 cantilever (64 000 voxels, 187 200 links), 30 steps: full-field comparison."""
    sc = scenarios.cantilever(40, 40, 40, tip_load=50.0)
    g, dt, _ = parity.run(product, sc, 30)
    o, _, _ = parity.run(oracle, sc, 30, dt=dt)
    err = parity.rel_errors(parity.snapshot(g), parity.snapshot(o), sc)
    assert err["pos"] <= 1e-9 and err["orient"] <= 1e-9, err


def test_full_size_properties_256(product):
    """BASELINE size (256^3, 16.7M voxels): size-independent properties instead of an oracle run.
    (1) symmetry: the load is in -z only and the lattice is mirror symmetric in y, so
        y-displacements of mirrored voxels are equal and opposite and x/z equal;
    (2) fixed face voxels never move; (3) total linear momentum change equals the applied
        impulse minus the reaction at the fixed face is not available without reactions, so
        check instead that momentum of free voxels in y sums to ~0."""
    n = 256
    sc = scenarios.cantilever(n, n, n, tip_load=1.0)
    s = scenarios.build(product, sc)
    dt = s.recommended_dt()
    assert s.step(dt, 10) is None
    pos = s.download("pos").reshape(n, n, n, 3)          # [k][j][i]
    nominal = sc.ijk.astype(np.float64).reshape(n, n, n, 3) * sc.voxel_size
    d = pos - nominal
    assert np.all(d[:, :, 0, :] == 0.0)                  # x = 0 face is fixed
    assert np.max(np.abs(d[:, :, -1, 2])) > 0            # the loaded face moved
    mirror = d[:, ::-1, :, :]
    # tolerance: 1e-9 of the displacement scale, but never below a few ulp of the coordinates
    tol = max(1e-9 * np.max(np.abs(d)), 16 * np.finfo(np.float64).eps * np.max(np.abs(pos)))
    assert np.max(np.abs(d[..., 0] - mirror[..., 0])) <= tol
    assert np.max(np.abs(d[..., 2] - mirror[..., 2])) <= tol
    assert np.max(np.abs(d[..., 1] + mirror[..., 1])) <= tol
    lm = s.download("linmom")
    assert abs(lm[:, 1].sum()) <= 1e-9 * np.abs(lm).sum()


@pytest.mark.parametrize("path", [7, 1, 0], ids=["lattice", "general", "auto-small"])
def test_state_info_reductions(product, oracle, path):
    """vx_state_info (device reductions) against the oracle's sequential float loops
    (CVoxelyze::stateInfo, src/Voxelyze.cpp:752-800); summation order differs, tolerance 1e-6 relative."""
    sc = scenarios.cantilever(10, 4, 3, tip_load=5.0)
    g, dt, _ = parity.run(product, sc, 400, path=path)
    o, _, _ = parity.run(oracle, sc, 400)
    for info in range(10):
        # signed quantities (stress, strain, pressure) cancel in TOTAL/AVERAGE: the reference's sequential float sum carries
        # up to n * eps * max|value| of rounding, so that is the scale the tolerance is tied to
        scale = max(abs(o.state_info(info, 0)), abs(o.state_info(info, 1)))
        n = g.n_links if info in (5, 6, 7) else g.n_voxels
        for typ in (0, 1, 2, 3):
            a, b = g.state_info(info, typ), o.state_info(info, typ)
            slack = (n if typ == 2 else 1) * 1.2e-7 * scale if typ >= 2 else 0.0
            assert abs(a - b) <= 1e-6 * max(abs(b), 1e-30) + slack + 1e-30, (info, typ, a, b, scale)
    # PRESSURE with different materials on the two ends of a link (strainRatio != 1) and Poisson's ratio in the denominator
    sc = cases.BY_NAME["mixed_six"].make()
    g, dt, _ = parity.run(product, sc, 300)
    o, _, _ = parity.run(oracle, sc, 300)
    scale = max(abs(o.state_info(8, 0)), abs(o.state_info(8, 1)))
    assert scale > 0
    for typ in (0, 1, 2, 3):
        a, b = g.state_info(8, typ), o.state_info(8, typ)
        assert abs(a - b) <= 1e-5 * scale, (typ, a, b, scale)


def test_collision_case_energy_and_centre_of_mass(product, oracle):
    """BASELINE north_star: cases with collisions are judged on energy and centre-of-mass trajectory.
    Plates case: kinetic energy within 1e-4 relative (or 1e-12 J absolute) and COM within 1e-9 m of the
    oracle at several checkpoints; watched pair sets identical at each checkpoint."""
    c = cases.BY_NAME["plates_16x4x2"]
    sc = c.make()
    g = scenarios.build(product, sc); o = scenarios.build(oracle, sc)
    dt = g.recommended_dt()
    assert dt == o.recommended_dt()
    for _ in range(6):
        g.step(dt, 500); o.step(dt, 500)
        ke_g, ke_o = g.state_info(2, 2), o.state_info(2, 2)
        assert abs(ke_g - ke_o) <= 1e-4 * abs(ke_o) + 1e-12, (ke_g, ke_o)
        com_g, com_o = g.download("pos").mean(axis=0), o.download("pos").mean(axis=0)
        assert np.max(np.abs(com_g - com_o)) <= 1e-9
        assert np.array_equal(g.collision_pairs(), o.collision_pairs())
        assert np.array_equal(g.download("linkflags") & 0xC, o.download("linkflags") & 0xC)


def test_split_asynchronous_steps_match_whole_steps_bitwise(product):
    """vx_step_begin / vx_step_enqueue(boundary, interior) / vx_step_end with the halo shipped on a second
    stream between the two parts (the multi-GPU overlap of SURVEY.md section 8e), emulated on ONE GPU:
    two z-slabs of a cantilever live in one process and trade pose planes by device copies."""
    import torch
    from voxelyze_b200 import slab
    nx, ny, nz, steps = 10, 9, 23, 37
    sc = scenarios.cantilever(nx, ny, nz, tip_load=25.0)
    whole, dt, _ = parity.run(product, sc, steps)
    main, comm = torch.cuda.current_stream(), torch.cuda.Stream()
    runs = [slab.SlabRunner(product, nx, ny, nz, r, 2, tip_load=25.0) for r in range(2)]
    for r in runs:
        assert r.sim.active_path() == 2
        r.sim.set_stream(main.cuda_stream)
    views = {}

    def plane(sim, z):
        p0, p1, n, rb = sim.pose_plane(z)
        for p in (p0, p1):
            if p not in views:
                views[p] = torch.as_tensor(slab._DevMem(p, n * rb), device="cuda")
        return views[p0], views[p1], n

    done = 0
    for chunk in (1, 20, steps - 21):                  # several calls: generations alternate across calls too
        for r in runs:
            r.sim.step_begin(dt)
        imported = None
        for _ in range(chunk):
            if imported is not None:
                main.wait_event(imported)
            for r in runs:
                r.sim.step_enqueue(r.sim.PART_Z_BOUNDARY)
            ready = torch.cuda.Event(); ready.record(main)
            for r in runs:
                r.sim.step_enqueue(r.sim.PART_Z_INTERIOR)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                for me, other in ((runs[0], runs[1]), (runs[1], runs[0])):
                    (peer, send_z, recv_z), = me._neighbours()
                    a0, a1, n = plane(other.sim, recv_z)         # the peer owns my ghost layer
                    b0, b1 = a0.clone(), a1.clone()              # the message
                    me.sim.halo_import(recv_z, b0.data_ptr(), b1.data_ptr(), n, comm.cuda_stream)
                imported = torch.cuda.Event(); imported.record(comm)
        main.wait_event(imported)
        for r in runs:
            assert r.sim.step_end() is None
        done += chunk
    assert done == steps
    for f in ("pos", "orient", "linmom", "angmom"):
        got = np.concatenate([r.owned_state(f) for r in runs])
        assert parity.bit_equal(got, whole.download(f)), f


def test_peer_memory_halo_steps_match_whole_steps_bitwise(product):
    """vx_peer_export/attach + vx_slab_step: each slab stores its boundary poses straight into the other's
    ghost layer from its own stream and spins on an arrival counter; here both slabs live in one process
    on one GPU (plain addresses instead of CUDA IPC mappings) and are stepped alternately."""
    from voxelyze_b200 import slab
    nx, ny, nz, steps = 9, 10, 21, 30
    sc = scenarios.cantilever(nx, ny, nz, tip_load=25.0)
    whole, dt, _ = parity.run(product, sc, steps)
    runs = [slab.SlabRunner(product, nx, ny, nz, r, 3, tip_load=25.0) for r in range(3)]
    assert all("GSKIP" in r.sim.kernel_name() for r in runs)     # the ghost planes at the slab ends carry no bricks
    slab.SlabRunner.connect_local(runs)
    for _ in range(steps):
        for r in runs:
            assert r.step(dt, 1) is None
    for f in ("pos", "orient", "linmom", "angmom"):
        got = np.concatenate([r.owned_state(f) for r in runs])
        assert parity.bit_equal(got, whole.download(f)), f
    # perturb one slab's boundary, re-publish with vx_slab_exchange on every slab, and keep matching
    first, n = runs[1].layer_index_range(runs[1].z0)
    pos = runs[1].sim.download("pos", first, n); pos[:, 2] += 1e-7
    runs[1].sim.upload("pos", pos, first)
    wfirst = runs[1].z0 * nx * ny
    wpos = whole.download("pos", wfirst, n); wpos[:, 2] += 1e-7
    whole.upload("pos", wpos, wfirst)
    for r in runs:
        r.exchange()
    whole.step(dt, 5)
    for _ in range(5):
        for r in runs:
            r.step(dt, 1)
    got = np.concatenate([r.owned_state("pos") for r in runs])
    assert parity.bit_equal(got, whole.download("pos"))


def test_any_scenario_runs_as_peer_memory_slabs_bitwise(product):
    """SlabRunner.from_scenario on the fused path with the in-kernel halo push: two materials with CTE, gravity, floor, a fixed
    face, forced voxels, a prescribed displacement, per-step ambient temperature (applied by the fused kernel on every slab) --
    three slabs of one process on one GPU against the unsplit run, bit for bit."""
    from voxelyze_b200 import slab
    from test_slab_gloo import _general_scenario
    sc = _general_scenario()
    whole = scenarios.build(product, sc, path=7); dt = whole.recommended_dt()
    runs = [slab.SlabRunner.from_scenario(product, sc, r, 3) for r in range(3)]
    assert all(r.sim.active_path() == 2 and "GSKIP" in r.sim.kernel_name() for r in runs) and whole.active_path() == 2
    slab.SlabRunner.connect_local(runs)
    for r in runs:
        r.exchange()                                     # initial temperature of the ghosts
    for k in range(120):
        t = 3.0 * np.sin(k / 10.0)
        whole.set_temperature_all(t)
        assert whole.step(dt, 1) is None
        for r in runs:
            r.set_temperature_all(t)
            assert r.step(dt, 1) is None
    index = np.concatenate([r.scenario_index[(r.z0 - r.lo) * r.plane:(r.z1 - r.lo) * r.plane] for r in runs])
    for f in ("pos", "orient", "linmom", "angmom", "temp", "voxflags"):
        got = np.concatenate([r.owned_state(f) for r in runs])
        assert parity.bit_equal(got, whole.download(f)[index]), f


@pytest.mark.parametrize("layout", [7, 0], ids=["fused", "auto"])
@pytest.mark.parametrize("case", ["cantilever", "plates", "robots"])
def test_checkpoint_resume_is_bit_identical(product, tmp_path, case, layout):
    """vx_save_state / vx_load_state: stop, restore into a freshly built handle, continue == never stopped.
    Lattice path (cantilever, ensemble with temperature) and general path with collisions and plastic links."""
    sc = {"cantilever": lambda: scenarios.cantilever(14, 5, 6, tip_load=40.0),
          "plates": lambda: scenarios.plate_stack(16, 8, 2, thick=3, gap=2, tip_load=1.0),
          "robots": lambda: scenarios.robot_ensemble(5, 4)}[case]()
    a = scenarios.build(product, sc, path=layout); dt = a.recommended_dt()
    if case == "robots":
        a.set_temperature_all(7.5)
    a.step(dt, 333)
    path = str(tmp_path / "state.bin")
    a.save_state(path)
    a.step(dt, 200)
    b = scenarios.build(product, sc, path=layout)
    b.load_state(path)
    assert b.time() == pytest.approx(333 * dt, rel=1e-4)
    b.step(dt, 200)
    sa, sb = parity.snapshot(a), parity.snapshot(b)
    for f in sa:
        assert parity.bit_equal(sa[f], sb[f]), f
    assert a.time() == b.time()
    # a file of another model is refused
    other = scenarios.build(product, scenarios.cantilever(14, 5, 7))
    with pytest.raises(capi.VxError):
        other.load_state(path)


@pytest.mark.parametrize("path", [7, 1, 0], ids=["lattice", "general", "auto-small"])
def test_link_state_upload_round_trip(product, path):
    """vx_download_link_state / vx_upload_link_state: a fresh handle that receives the voxel and link state of a
    running one continues bit-identically (what the facade does across setVoxel edits)."""
    sc = scenarios.cantilever(12, 4, 3, tip_load=60.0)
    a = scenarios.build(product, sc, path=path); dt = a.recommended_dt()
    a.step(dt, 400)
    b = scenarios.build(product, sc, path=path)
    b.step(dt, 1)                          # so that both handles damp their next step with the same previous dt
    for f in ("pos", "orient", "linmom", "angmom", "temp"):
        b.upload(f, a.download(f))
    rec = a.download_link_state()
    assert (rec["flags"] & 2).all() and np.abs(rec["pos2"]).max() > 0
    b.upload_link_state(rec)
    back = b.download_link_state()
    for name in ("pos2", "angle1v", "angle2v", "strain", "max_strain", "strain_offset", "stress", "flags"):
        assert np.array_equal(back[name], rec[name]), name
    a.step(dt, 150); b.step(dt, 150)
    for f in ("pos", "orient", "linmom", "angmom"):
        assert parity.bit_equal(a.download(f), b.download(f)), f
    sa, sb = a.download_link_state(), b.download_link_state()
    assert sa.tobytes() == sb.tobytes()


def test_enabling_collisions_mid_run_matches_the_oracle(product, oracle):
    """CVoxelyze::enableCollisions may be called at any time (src/Voxelyze.cpp:612-622); on the device the model is laid
    out again (collision tables) while every voxel and link keeps its state."""
    sc = scenarios.drop_block(6)
    g = scenarios.build(product, sc); o = scenarios.build(oracle, sc)
    assert g.active_path() == 1                       # 216 voxels, no collisions yet: small-model kernel on the general layout
    dt = g.recommended_dt()
    g.step(dt, 300); o.step(dt, 300)
    g.enable_collisions(True); o.enable_collisions(True)
    assert g.active_path() == 2                       # the fused kernels gather contact forces themselves
    assert abs(g.time() - o.time()) <= 1e-9
    g.step(dt, 500); o.step(dt, 500)
    err = parity.rel_errors(parity.snapshot(g), parity.snapshot(o), sc)
    assert err["pos"] <= 1e-7 and err["orient"] <= 1e-7, err
    assert np.array_equal(g.collision_pairs(), o.collision_pairs())


@pytest.mark.parametrize("path", [5, 7], ids=["cpasync", "tma"])
def test_ensemble_members_stay_independent_on_both_stagings(product, path):
    """Ensemble members are stacked along z in the device arrays; a brick at the top of one member sees the next member's
    bottom planes in its +Z face (and, with TMA boxes, in its own box when nz is odd).  Odd box edges, several materials,
    floor contact and temperature: both staging flavours must equal the general path bit for bit."""
    sc = scenarios.robot_ensemble(7, 5)                  # 5 x 5 x 5 robots: partial bricks on every axis
    snaps = {}
    for p in (1, path):
        sim = scenarios.build(product, sc, path=p); dt = sim.recommended_dt()
        assert sim.active_path() == (1 if p == 1 else 2)
        for k in range(60):
            sim.set_temperature_all(scenarios.robot_temperature(k * dt) * 3)
            sim.step(dt, 5)
        snaps[p] = parity.snapshot(sim)
    for f in snaps[1]:
        assert parity.bit_equal(snaps[1][f], snaps[path][f]), f


def test_single_material_grids_equal_the_general_path_bitwise(product):
    """The single-material instantiation (k_lattice_tma<UNI>: no table look-ups) on the three grids it runs on -- a box with odd
    edges on every axis, an ensemble stacked along z, and the brick-group list of a sparse body -- against the general
    two-kernel path, bit for bit, with both stagings."""
    M = Material(E=2e6, rho=1.2e3, zeta_global=0.02, zeta_internal=0.5, cte=0.004, mu_static=1.0, mu_kinetic=0.5)
    # (a) odd box
    sc = scenarios.cantilever(21, 14, 11, tip_load=40.0)
    # (b) 7 members of 9 x 6 x 5 on a floor, different loads, ambient temperature program
    base = scenarios.box_ijk(9, 6, 5)
    ijk = np.tile(base, (7, 1)); sid = np.repeat(np.arange(7), len(base)).astype(np.int32)
    ens = scenarios.Scenario("uni_members", 0.005, [M], ijk, np.zeros(len(ijk), np.uint16), sim_id=sid, gravity=1.0, floor=True)
    top = np.nonzero((ijk[:, 2] == 4) & (ijk[:, 0] == 8))[0]
    ens.ext_voxel = top.astype(np.int32); ens.ext_dof = np.zeros(len(top), np.uint8)
    ens.ext_force = np.stack([0.002 * (sid[top] - 3), np.zeros(len(top)), -0.001 * sid[top]], 1).astype(np.float32)
    # (c) an L of one material
    lijk = np.array([[i, j, k] for k in range(8) for j in range(24) for i in range(24) if j < 8 or i < 8], np.int32)
    ell = scenarios.Scenario("ell_uni", 0.005, [M], lijk, np.zeros(len(lijk), np.uint16), gravity=0.3)
    fixed = np.nonzero(lijk[:, 0] == 23)[0]
    ell.ext_voxel = fixed.astype(np.int32); ell.ext_dof = np.full(len(fixed), 0x3F, np.uint8); ell.ext_force = np.zeros((len(fixed), 3), np.float32)
    for name, s, path, steps, temp in (("box", sc, 5, 300, False), ("members", ens, 5, 300, True), ("ell", ell, 0, 300, False)):
        snaps = {}
        for p in (1, 7, path):
            sim = scenarios.build(product, s, path=p); dt = sim.recommended_dt()
            assert sim.active_path() == (1 if p == 1 else 2)
            for k in range(steps // 5):
                if temp:
                    sim.set_temperature_all(scenarios.robot_temperature(k * dt) * 3)
                assert sim.step(dt, 5) is None
            snaps[p] = parity.snapshot(sim)
        for p in (7, path):
            for f in snaps[1]:
                assert parity.bit_equal(snaps[1][f], snaps[p][f]), (name, p, f)


def test_sparse_l_shaped_body_runs_fused_from_a_brick_group_list(product, oracle):
    """A body that fills 56 % of its bounding box (two arms of an L, 24 x 8 x 8 and 8 x 24 x 8): padded to the box, but only the
    occupied 8 x 8 x 4 brick groups are launched (LatFrame::groups; 10 of 18 here).  Same bits as the general path, parity
    with the oracle, and the caller sees only its own voxels."""
    ijk = np.array([[i, j, k] for k in range(8) for j in range(24) for i in range(24) if j < 8 or i < 8], np.int32)
    assert len(ijk) == 2560
    mats = [Material(E=1e6, rho=1e3, zeta_global=0.01), Material(E=2e6, rho=1.5e3, zeta_global=0.01)]
    mat = ((ijk[:, 0] // 3 + ijk[:, 1] // 3) % 2).astype(np.uint16)
    sc = scenarios.Scenario("ell_3d", 0.005, mats, ijk, mat, gravity=0.2)
    fixed = np.nonzero(ijk[:, 0] == 23)[0]
    load = np.nonzero(ijk[:, 1] == 23)[0]
    sc.ext_voxel = np.concatenate([fixed, load]).astype(np.int32)
    sc.ext_dof = np.concatenate([np.full(len(fixed), 0x3F), np.zeros(len(load))]).astype(np.uint8)
    f = np.zeros((len(sc.ext_voxel), 3), np.float32); f[len(fixed):] = [0.001, 0.0, -0.002]
    sc.ext_force = f
    runs = {}
    for path in (0, 1):
        g = scenarios.build(product, sc, path=path); dt = g.recommended_dt()
        assert g.active_path() == (1 if path == 1 else 2) and g.n_voxels == len(ijk)
        assert g.step(dt, 400) is None
        runs[path] = g
    assert "k_lattice_tma" in runs[0].kernel_name()
    a, c = parity.snapshot(runs[0]), parity.snapshot(runs[1])
    for fld in a:
        assert parity.bit_equal(a[fld], c[fld]), fld
    o, _, _ = parity.run(oracle, sc, 400, dt=dt)
    err = parity.rel_errors(a, parity.snapshot(o), sc)
    assert err["pos"] <= 1e-9 and err["orient"] <= 1e-9, err


def test_box_with_holes_runs_fused_and_matches_the_general_path_bitwise(product, oracle):
    """An irregular body (a block with a notch, a pocket and a through hole; two materials) is padded with inert fill
    cells to its bounding box and stepped by the fused lattice kernels: same bits as the general path, same voxel and
    link lists for the caller, stateInfo over real voxels only, and the usual parity with the oracle."""
    ijk = [[i, j, k] for k in range(5) for j in range(6) for i in range(9)
           if not (i >= 6 and k >= 3) and not (2 <= i <= 3 and 2 <= j <= 3) and not (i == 7 and j == 1 and k <= 1)]
    ijk = np.array(ijk, np.int32)
    mats = [Material(E=1e6, rho=1e3, zeta_global=0.02), Material(E=4e6, rho=2e3, zeta_global=0.02)]
    mat = ((ijk[:, 0] + ijk[:, 2]) % 2).astype(np.uint16)
    sc = scenarios.Scenario("holes", 0.002, mats, ijk, mat, gravity=1.0, floor=True)
    fixed = np.nonzero(ijk[:, 0] == 0)[0]; load = np.nonzero(ijk[:, 0] == 8)[0]
    sc = scenarios._externals(sc, fixed, load, [0.0, 0.002, -0.004])
    assert len(ijk) < 9 * 6 * 5 and 9 * 6 * 5 <= 1.6 * len(ijk)
    runs = {}
    for path in (0, 5, 7, 1):
        g = scenarios.build(product, sc, path=path); dt = g.recommended_dt()
        assert g.active_path() == parity.layout(path, len(ijk)) and g.n_voxels == len(ijk)
        g.step(dt, 1200)
        runs[path] = g
    o = scenarios.build(oracle, sc); o.step(dt, 1200)
    a, c = parity.snapshot(runs[0]), parity.snapshot(runs[1])
    for path in (0, 5, 7):
        b = parity.snapshot(runs[path])
        for f in b:
            assert parity.bit_equal(b[f], c[f]), (path, f)
    assert np.array_equal(np.stack(runs[0].links()), np.stack(o.links()))
    err = parity.rel_errors(a, parity.snapshot(o), sc)
    assert err["pos"] <= 1e-9 and err["orient"] <= 1e-9, err
    for info in (0, 2, 6, 8, 9):
        for typ in (0, 1, 3):
            x, y = runs[0].state_info(info, typ), o.state_info(info, typ)
            scale = max(abs(o.state_info(info, 0)), abs(o.state_info(info, 1)))
            assert abs(x - y) <= 1e-5 * scale + 1e-30, (info, typ, x, y)


def test_collisions_on_the_fused_path_match_the_general_path_bitwise(product):
    """Plates bending onto each other (self-collisions, two bilinear materials, gaps between the plates = a box with holes):
    the fused lattice kernels with their contact-force gather against the general path, bit for bit, pair sets included."""
    sc = scenarios.plate_stack(16, 4, 2, 3, 2, tip_load=0.5)
    runs = {}
    for path in (0, 5, 1):
        g = scenarios.build(product, sc, path=path); dt = g.recommended_dt()
        assert g.active_path() == (1 if path == 1 else 2)
        g.step(dt, 2500)
        runs[path] = g
    assert len(runs[1].collision_pairs()) > 0
    for path in (0, 5):
        a, c = parity.snapshot(runs[path]), parity.snapshot(runs[1])
        for f in a:
            assert parity.bit_equal(a[f], c[f]), (path, f)
        assert np.array_equal(runs[path].collision_pairs(), runs[1].collision_pairs())


@pytest.mark.parametrize("path", [7, 0], ids=["fused", "auto-small"])
def test_poissons_ratio_switched_on_mid_run_keeps_the_state(product, oracle, path):
    """The reference lets a caller change Poisson's ratio at any time (src/VX_Link.cpp:160-166 'catches when we disable
    poissons mid-simulation').  A model on the fused layout stays there: its per-voxel Poisson strains are created from the
    current link strains when nu becomes non-zero; voxel and link state, time and previousDt go on, and later calls (gravity,
    more steps) keep working."""
    import copy

    def run(lib):
        sc = scenarios.cantilever(8, 3, 3, tip_load=20.0)
        s = scenarios.build(lib, sc, path=path); dt = s.recommended_dt()
        s.step(dt, 150)
        m = copy.copy(sc.materials[0]); m.nu = 0.3
        s.set_materials([m])
        dt2 = s.recommended_dt()
        assert s.step(dt2 * 0.5, 150) is None
        s.set_gravity(0.5)                                  # every later table upload must still work
        assert s.step(dt2 * 0.5, 50) is None
        return s, sc, dt, dt2

    (g, sc, dtg, dtg2), (o, _, dto, dto2) = run(product), run(oracle)
    assert g.active_path() == parity.layout(path, sc.n_voxels) and dtg == dto
    assert abs(dtg2 - dto2) <= 1e-6 * dto2
    err = parity.rel_errors(parity.snapshot(g), parity.snapshot(o), sc)
    assert err["pos"] <= 1e-6 and err["orient"] <= 1e-6, err
    assert abs(g.time() - o.time()) <= 1e-6 * o.time()
