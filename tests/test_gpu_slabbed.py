"""vx_slabbed_* on the CUDA library: the whole model in the caller's numbering, z-slabs on the devices of ONE process, halo by
the step kernels' own peer stores.  Bit for bit against the unsplit run.  With one GPU all slabs share it (lock-step stepping);
with two or more (gpurun --gpus 2) every slab has its own device and all of them are queued before any is waited for."""
import numpy as np
import pytest

import parity
from test_slab_gloo import _general_scenario
from test_slabbed import check_slabbed_against_whole, check_state_edits, holes_scenario, poisson_scenario, VOXEL_FIELDS, LINK_FIELDS
from voxelyze_b200 import capi, scenarios

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_slabbed_handle_runs_on_peer_stores_bitwise(product):
    whole, multi, dt = check_slabbed_against_whole(product, _general_scenario(), [0, 0, 0], 120, expect_halo=2)
    assert all(multi.slab(k).active_path() == 2 and "GSKIP" in multi.slab(k).kernel_name() for k in range(3))
    assert multi.launch_count() > 0


def test_slabbed_body_with_holes(product):
    whole, multi, dt = check_slabbed_against_whole(product, holes_scenario(), [0] * 4, 150, temperature_program=False, chunk=10)
    assert multi.halo_mode in (1, 2)


def test_slabbed_poisson_materials_on_peer_stores(product):
    """Poisson coupling across the cuts on the fused kernel: k_lattice_tma<.., PUSH, POISSON> stores a boundary voxel's new Poisson
    strain into the neighbour's ghost plane next to its pose."""
    sc = poisson_scenario()
    whole, multi, dt = check_slabbed_against_whole(product, sc, [0, 0, 0], 150, temperature_program=False, expect_halo=2, chunk=5, path=7)
    assert all(multi.slab(k).active_path() == 2 for k in range(3)) and whole.active_path() == 2
    assert np.abs(whole.download("pstrain")).max() > 1e-4
    assert parity.bit_equal(multi.download("pstrain"), whole.download("pstrain"))
    assert whole.step(-1.0, 7) is None and multi.step(-1.0, 7) is None       # the stable step re-evaluated before every step
    assert whole.time() == multi.time()
    for f in VOXEL_FIELDS + LINK_FIELDS:
        assert parity.bit_equal(multi.download(f), whole.download(f)), f
    sc0 = poisson_scenario(); sc0.materials[0].nu = 0.0           # switched on mid-run
    whole, multi, dt = check_slabbed_against_whole(product, sc0, [0, 0, 0], 60, temperature_program=False, chunk=5, path=7)
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


def test_slabbed_state_edits_reach_every_copy(product):
    check_state_edits(product, [0, 0, 0])


def test_slabbed_larger_lattice_many_steps_per_call(product):
    sc = scenarios.cantilever(40, 24, 48, tip_load=60.0)
    check_slabbed_against_whole(product, sc, [0, 0], 200, temperature_program=False, expect_halo=2, chunk=50)


def test_slabbed_divergence(product):
    sc = scenarios.cantilever(12, 6, 16, tip_load=1.0)
    whole, multi = scenarios.build(product, sc, path=7), scenarios.build_slabbed(product, sc, [0, 0])
    dt = 40.0 * whole.recommended_dt()
    a, b = whole.step(dt, 400), multi.step(dt, 400)
    assert a is not None and a == b


@pytest.mark.parametrize("n_dev", [2, 4])
def test_slabbed_one_slab_per_device(product, n_dev):
    """Distinct devices: vx_slab_step_begin on every slab, then vx_slab_step_finish -- no host work per step."""
    if _gpus() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs (gpurun --gpus {n_dev})")
    sc = scenarios.cantilever(64, 32, 64, tip_load=200.0)
    whole, multi, dt = check_slabbed_against_whole(product, sc, list(range(n_dev)), 300, temperature_program=False, expect_halo=2, chunk=100)
    multi.reset(); whole.reset()
    assert multi.step(dt, 37) is None and whole.step(dt, 37) is None
    for f in VOXEL_FIELDS:
        assert parity.bit_equal(multi.download(f), whole.download(f)), f


def test_state_moves_between_handles_with_its_clock(product):
    """What CVoxelyze::setDevices does mid-run: voxel fields, link records and vx_set_clock (time, CVX_Voxel::previousDt) into a
    fresh handle -- one device to three slabs and back -- and the run goes on with the bits of the run that never moved."""
    sc = _general_scenario()
    a = scenarios.build(product, sc, path=7); dt = a.recommended_dt()
    assert a.step(dt, 80) is None
    b = scenarios.build_slabbed(product, sc, [0, 0, 0])
    c = scenarios.build(product, sc, path=7)
    src = a
    for dst in (b, c):
        for f in ("pos", "orient", "linmom", "angmom", "temp", "voxflags"):
            dst.upload(f, src.download(f))
        dst.upload_link_state(src.download_link_state())
        dst.set_clock(src.time(), dt)
        assert dst.time() == src.time()
        assert a.step(dt, 40) is None and dst.step(dt, 40) is None
        for f in VOXEL_FIELDS:
            assert parity.bit_equal(dst.download(f), a.download(f)), (f, type(dst).__name__)
        src = dst


def test_slabbed_poisson_one_slab_per_device(product):
    """Poisson strains through the peer stores between DEVICES: all slabs queued (n steps each) before any is waited for."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    n_dev = min(_gpus(), 4)
    sc = poisson_scenario()
    whole, multi, dt = check_slabbed_against_whole(product, sc, list(range(n_dev)), 200, temperature_program=False, expect_halo=2, chunk=50, path=7)
    assert parity.bit_equal(multi.download("pstrain"), whole.download("pstrain"))


def test_slabbed_checkpoint_resume_is_bit_identical(product, tmp_path):
    """vx_slabbed_save_state / load_state: one file per slab; a freshly built slabbed handle continues with the bits of the run
    that never stopped (ghost planes and both generations are in the files: no exchange needed)."""
    sc = _general_scenario()
    a = scenarios.build_slabbed(product, sc, [0, 0, 0]); dt = a.recommended_dt()
    assert a.step(dt, 77) is None
    path = str(tmp_path / "slabbed_state")
    a.save_state(path)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["slabbed_state.0of3", "slabbed_state.1of3", "slabbed_state.2of3"]
    assert a.step(dt, 50) is None
    b = scenarios.build_slabbed(product, sc, [0, 0, 0])
    b.load_state(path)
    assert b.time() == pytest.approx(77 * dt, rel=1e-4) and b.step(dt, 50) is None
    for f in VOXEL_FIELDS + LINK_FIELDS:
        assert parity.bit_equal(a.download(f), b.download(f)), f
    c = scenarios.build_slabbed(product, sc, [0, 0])             # another cut: refused (no such files)
    with pytest.raises(capi.VxError):
        c.load_state(path)


def test_slabbed_model_cut_again_with_another_slab_count(product):
    """vx_slabbed_set_voxels on a handle that has been stepping: a thin model on two of three slabs, then a tall one on all three --
    the slab that sat idle joins with a fresh exchange count."""
    thin, tall = scenarios.cantilever(8, 4, 5, tip_load=5.0), scenarios.cantilever(8, 4, 14, tip_load=5.0)
    m = scenarios.build_slabbed(product, thin, [0, 0, 0])
    assert m.n_slabs == 2 and m.step(m.recommended_dt(), 25) is None
    m.set_voxels(tall.ijk, tall.mat)
    m.set_externals(tall.ext_voxel, tall.ext_dof, tall.ext_force)
    assert m.n_slabs == 3 and m.halo_mode == 2
    whole = scenarios.build(product, tall, path=7); dt = whole.recommended_dt()
    assert m.step(dt, 60) is None and whole.step(dt, 60) is None
    for f in VOXEL_FIELDS:
        assert parity.bit_equal(m.download(f), whole.download(f)), f
