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
