"""N > 1 host logic on CPU: world_size-2 gloo run of the z-slab decomposition.

The slab driver (voxelyze_b200/slab.py) is transport-agnostic; here it runs over gloo with the
oracle library and host exchange, and the stitched result must be BIT-IDENTICAL to the unsplit
run: the halo only moves data, cut-crossing links are evaluated redundantly from identical
inputs.  (On the GPU the same class runs over NCCL with device-side exchange.)"""
import os
import socket

import numpy as np
import pytest

import parity
from voxelyze_b200 import capi, scenarios, slab

NX, NY, NZ, STEPS = 6, 3, 7, 120


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = capi.load_oracle()
    r = slab.SlabRunner(lib, NX, NY, NZ, rank, world, voxel_size=0.005, tip_load=1.0, host_exchange=True)
    dt = r.recommended_dt()
    r.step(dt, STEPS)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), pos=r.owned_state("pos"), orient=r.owned_state("orient"),
             linmom=r.owned_state("linmom"), z0=r.z0, z1=r.z1, dt=dt)
    dist.barrier()
    dist.destroy_process_group()


def test_slab_ranges_cover_the_lattice():
    for nz in (7, 8, 512):
        for world in (1, 2, 3, 4, 8):
            r = [slab.slab_range(nz, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nz
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


@pytest.mark.parametrize("world", [2, 3])
def test_two_slabs_equal_unsplit_run_bitwise(built, tmp_path, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    sc = scenarios.cantilever(NX, NY, NZ, voxel_size=0.005, tip_load=1.0)
    whole, dt, _ = parity.run(capi.load_oracle(), sc, STEPS)
    assert np.float32(dt) == np.float32(parts[0]["dt"])
    for f in ("pos", "orient", "linmom"):
        stitched = np.concatenate([p[f] for p in parts])
        assert np.array_equal(stitched, whole.download(f)), f


def _worker_scenario(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = slab.SlabRunner.from_scenario(capi.load_oracle(), _general_scenario(), rank, world, host_exchange=True)
    dt = r.recommended_dt()
    div = None
    for k in range(90):
        r.set_temperature_all(3.0 * np.sin(k / 10.0))
        div = r.step(dt, 1, check_divergence=True) if div is None else div
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), pos=r.owned_state("pos"), orient=r.owned_state("orient"), temp=r.owned_state("temp"),
             index=r.scenario_index[(r.z0 - r.lo) * r.plane:(r.z1 - r.lo) * r.plane], dt=dt, div=-1 if div is None else div)
    dist.barrier()
    dist.destroy_process_group()


def _general_scenario():
    """Two materials with different CTE, gravity + floor, a fixed face, forced voxels and a prescribed displacement, lattice
    indices that do not start at zero, setVoxel order NOT plane-major: everything from_scenario has to carry."""
    from voxelyze_b200.capi import Material, DOF_ALL
    ijk = scenarios.box_ijk(5, 3, 9, origin=(2, -1, 1))
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(ijk))
    ijk = ijk[perm]
    mats = [Material(E=1e6, rho=1e3, cte=0.01, zeta_global=0.02, mu_static=1.0, mu_kinetic=0.5), Material(E=3e6, rho=2e3, cte=-0.005, zeta_global=0.02)]
    mat = ((ijk[:, 0] + ijk[:, 2]) & 1).astype(np.uint16)
    sc = scenarios.Scenario("slab_general", 0.005, mats, ijk, mat, gravity=0.3, floor=True, temperature=1.5)
    fixed = np.nonzero(ijk[:, 0] == 2)[0]
    load = np.nonzero(ijk[:, 0] == 6)[0]
    moved = np.nonzero((ijk[:, 0] == 4) & (ijk[:, 1] == 0) & (ijk[:, 2] == 5))[0]
    sc.ext_voxel = np.concatenate([fixed, load, moved]).astype(np.int32)
    sc.ext_dof = np.concatenate([np.full(len(fixed), DOF_ALL), np.zeros(len(load)), np.full(len(moved), 0x07)]).astype(np.uint8)
    f = np.zeros((len(sc.ext_voxel), 3), np.float32); f[len(fixed):len(fixed) + len(load)] = [0.0, 0.002, -0.004]
    t = np.zeros((len(sc.ext_voxel), 3), np.float64); t[-len(moved):] = [0.0, 0.0, 1e-4]
    sc.ext_force, sc.ext_translation = f, t
    return sc


def test_any_scenario_splits_into_slabs_bitwise(built, tmp_path):
    """SlabRunner.from_scenario: materials, externals, gravity, floor, initial and per-step temperature of an arbitrary full
    box, three slabs over gloo, against the unsplit run of the same calls."""
    import torch.multiprocessing as mp
    world = 3
    mp.spawn(_worker_scenario, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    sc = _general_scenario()
    whole = scenarios.build(capi.load_oracle(), sc)
    dt = whole.recommended_dt()
    assert np.float32(dt) == np.float32(parts[0]["dt"])
    for k in range(90):
        whole.set_temperature_all(3.0 * np.sin(k / 10.0))
        assert whole.step(dt, 1) is None
    index = np.concatenate([p["index"] for p in parts])
    assert sorted(index.tolist()) == list(range(sc.n_voxels)) and all(int(p["div"]) == -1 for p in parts)
    for f in ("pos", "orient", "temp"):
        stitched = np.concatenate([p[f] for p in parts])
        assert np.array_equal(stitched, whole.download(f)[index]), f


def _worker_poisson(rank, world, port, out_dir):
    import torch.distributed as dist
    from test_slabbed import poisson_scenario
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = slab.SlabRunner.from_scenario(capi.load_oracle(), poisson_scenario(), rank, world, host_exchange=True)
    dt = r.recommended_dt()
    assert r.poisson and r.step(dt, 80) is None
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), pos=r.owned_state("pos"), orient=r.owned_state("orient"), pstrain=r.owned_state("pstrain"),
             index=r.scenario_index[(r.z0 - r.lo) * r.plane:(r.z1 - r.lo) * r.plane], dt=dt)
    dist.barrier()
    dist.destroy_process_group()


def test_poisson_materials_split_into_slabs_bitwise(built, tmp_path):
    """Poisson coupling across the cuts: the ghosts' Poisson strains travel with the halo (third field of the host exchange)."""
    import torch.multiprocessing as mp
    from test_slabbed import poisson_scenario
    world = 2
    mp.spawn(_worker_poisson, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    whole = scenarios.build(capi.load_oracle(), poisson_scenario())
    dt = whole.recommended_dt()
    assert np.float32(dt) == np.float32(parts[0]["dt"]) and whole.step(dt, 80) is None
    index = np.concatenate([p["index"] for p in parts])
    for f in ("pos", "orient", "pstrain"):
        stitched = np.concatenate([p[f] for p in parts])
        assert np.array_equal(stitched, whole.download(f)[index]), f


def _worker_ensemble(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    robots = 6 // world                                   # bench.py --config c4: rank r holds robots [r * robots, (r + 1) * robots)
    sc = scenarios.robot_ensemble(robots, 4, first_seed=rank * robots)
    r = slab.EnsembleRunner(scenarios.build(capi.load_oracle(), sc), world)
    dt = r.recommended_dt()
    assert r.step(dt, 1) is None and r.step(dt, 59) is None          # the per-step calls of the e2e leg, then the program in one call
    nv, nl = r.global_counts()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), pos=r.sim.download("pos"), orient=r.sim.download("orient"), temp=r.sim.download("temp"), dt=dt, nv=nv, nl=nl)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ensemble_shards_equal_the_whole_population_bitwise(built, tmp_path, world):
    """BASELINE config C4 sharded like bench.py --config c4 does (no data-path communication, "replicas only"): the shards'
    robots, stepped with the temperature program handed over in one call (vx_step_ambient), equal the same robots in one handle
    stepped with setAmbientTemperature + doTimeStep in turn."""
    import torch.multiprocessing as mp
    mp.spawn(_worker_ensemble, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    sc = scenarios.robot_ensemble(6, 4)
    whole = scenarios.build(capi.load_oracle(), sc); dt = whole.recommended_dt()
    assert np.float32(dt) == np.float32(parts[0]["dt"])
    assert int(parts[0]["nv"]) == whole.n_voxels and int(parts[0]["nl"]) == whole.n_links
    for _ in range(60):
        whole.set_temperature_all(scenarios.robot_temperature(whole.time()))
        assert whole.step(dt, 1) is None
    for f in ("pos", "orient", "temp"):
        assert np.array_equal(np.concatenate([p[f] for p in parts]), whole.download(f)), f
