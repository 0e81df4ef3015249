"""z-slab decomposition on real GPUs: 2 ranks over NCCL (device-side halo exchange of the pose
planes) must reproduce the single-GPU run of the same lattice BIT FOR BIT - the halo only moves
data and cut-crossing links are evaluated redundantly from identical inputs."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["VX_ROOT"])
from voxelyze_b200 import capi, slab
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
path = int(os.environ["VX_PATH"])
r = slab.SlabRunner(capi.load_product(), 24, 10, 18, rank, world, device=rank, tip_load=40.0, path=path,
                    overlap=os.environ["VX_OVERLAP"] != "0", peer=os.environ["VX_OVERLAP"] == "2")
assert r.overlap == (os.environ["VX_OVERLAP"] != "0")
assert r.peer == (os.environ["VX_OVERLAP"] == "2"), getattr(r, "peer_error", "")
r.sim.set_stream(torch.cuda.current_stream().cuda_stream)
dt = r.recommended_dt()
r.step(dt, 60)
torch.cuda.synchronize()
r.step(dt, 1); r.step(dt, 3)          # a second and third call: generations keep alternating across calls
torch.cuda.synchronize()
np.savez(os.path.join(os.environ["VX_OUT"], f"rank{rank}_p{path}.npz"), pos=r.owned_state("pos"), orient=r.owned_state("orient"),
         linmom=r.owned_state("linmom"), dt=dt, active=r.sim.active_path())
dist.barrier(); dist.destroy_process_group()
'''


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("path,overlap", [(0, 2), (0, 1), (0, 0), (1, 0)], ids=["lattice-peer-memory", "lattice-nccl-overlapped", "lattice-nccl-serial", "general"])
def test_two_gpu_slabs_equal_single_gpu_bitwise(product, tmp_path, path, overlap):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, VX_ROOT=ROOT, VX_OUT=str(tmp_path), VX_PATH=str(path), VX_OVERLAP=str(overlap))
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                    "--master-port", str(port), str(script)], check=True, env=env, timeout=600)
    parts = [np.load(tmp_path / f"rank{k}_p{path}.npz") for k in range(2)]
    from voxelyze_b200 import scenarios
    import parity
    sc = scenarios.cantilever(24, 10, 18, tip_load=40.0)
    whole, dt, _ = parity.run(product, sc, 64, path=path)
    assert np.float32(dt) == np.float32(parts[0]["dt"])
    assert int(parts[0]["active"]) == (2 if path == 0 else 1)
    for f in ("pos", "orient", "linmom"):
        got, want = np.concatenate([p[f] for p in parts]), whole.download(f)
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, (f, bad.size, bad[:8], sc.ijk[bad[:8]].tolist(), float(np.abs(got - want).max()))
