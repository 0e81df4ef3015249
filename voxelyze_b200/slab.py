"""z-slab domain decomposition of one large lattice (BASELINE config C5b).

One process per GPU.  Rank r owns the z-layers [z0, z1) plus one ghost layer on each cut.
A voxel update needs only the *old* poses of its face neighbours, so the only exchange step
is: after every step each rank sends the pose (position + orientation, 56 B) of its two
boundary layers to its z-neighbours, which store them into their ghost layers.  Links that
cross a cut are evaluated redundantly on both sides from identical inputs (identical bits),
so no force ever travels.  Transport: NCCL send/recv over NVLink via torch.distributed on
views of the library's own device arrays (no staging copy on the send side).

The same class runs on CPU tensors over gloo with any C-ABI implementation (host exchange
through vx_download / vx_upload); tests use that to check the partition logic bit for bit
against an unsplit run.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import scenarios
from .capi import DOF_ALL, Material, Sim, VxLib, VF_GHOST


def slab_range(nz: int, rank: int, world: int) -> Tuple[int, int]:
    """Owned z-layers [z0, z1) of `rank`; layers are dealt as evenly as possible."""
    base, rem = divmod(nz, world)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def _kernel_name(sim: Sim) -> str:
    return sim.kernel_name()


def _path_name(sim: Sim) -> str:
    return "fused lattice path" if sim.active_path() == 2 else "general two-kernel path"


class _DevMem:
    """Exposes a raw device allocation to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class SingleRunner:
    """Whole lattice on one GPU; same interface as SlabRunner."""

    def __init__(self, sim: Sim):
        self.sim = sim

    def recommended_dt(self) -> float:
        return self.sim.recommended_dt()

    def step(self, dt: float, n: int):
        return self.sim.step(dt, n)

    def step_profile(self, dt: float, n: int):
        return self.sim.step_profile(dt, n)

    def global_counts(self):
        return self.sim.n_voxels, self.sim.n_links

    def local_counts(self):
        return self.sim.n_voxels, self.sim.n_links

    def read_probe(self):
        return self.sim.download("pos", self.sim.n_voxels - 1, 1)

    def dominant_kernel(self) -> str:
        return _kernel_name(self.sim)

    def path_name(self) -> str:
        return _path_name(self.sim)

    def h2d_bytes_per_step(self) -> int:
        return 4                                    # dt

    def close(self):
        self.sim.close()


class EnsembleRunner(SingleRunner):
    """C4: this rank's share of a population of independent robots in ONE handle (members never interact), the
    ambient temperature program of SURVEY.md section 8d set from the host before EVERY step like a caller of
    CVoxelyze::setAmbientTemperature would; `world` ranks hold equal shares and never communicate."""

    def __init__(self, sim: Sim, world: int = 1):
        super().__init__(sim)
        self.world = world

    def _advance(self, dt: float):
        self.sim.set_temperature_all(scenarios.robot_temperature(self.sim.time()))

    def step(self, dt: float, n: int):
        """n == 1: the two calls a user of the class API makes per step (setAmbientTemperature, doTimeStep -- bench.py's e2e leg);
        n > 1: the same program handed over in one call (vx_step_ambient: one launch per step, no host round trip in between).
        The device accumulates the simulated time in float (src/Voxelyze.cpp:282); so does the program here, so both ways
        set the same temperatures."""
        if n == 1:
            self._advance(dt)
            return self.sim.step(dt, 1)
        t, temps = np.float32(self.sim.time()), []
        for _ in range(n):
            temps.append(scenarios.robot_temperature(float(t)))
            t = np.float32(t + np.float32(dt))
        return self.sim.step_ambient(dt, temps)

    def step_profile(self, dt: float, n: int):
        tot, launches = None, None
        for _ in range(n):
            self._advance(dt)
            ms, ln = self.sim.step_profile(dt, 1)
            tot = ms if tot is None else {k: tot[k] + ms[k] for k in ms}
            launches = ln if launches is None else [a + b for a, b in zip(launches, ln)]
        return tot, launches

    def global_counts(self):
        return self.sim.n_voxels * self.world, self.sim.n_links * self.world

    def h2d_bytes_per_step(self) -> int:
        return 8                                    # dt + ambient temperature

    def path_name(self) -> str:
        return f"{_path_name(self.sim)}, {self.sim.n_voxels // 1000} robots on this GPU"


class SlabRunner:
    """One rank's z-slab of a single lattice.  `SlabRunner(lib, nx, ny, nz, ...)` builds the cantilever pattern of C5 (x=0 face
    fixed, -z load on the x=nx-1 face) without ever materialising the whole lattice on a host; `SlabRunner.from_scenario`
    splits ANY full-box scenario -- several materials, externals (fixed / forced / prescribed voxels), gravity, floor,
    temperature -- so that whatever runs on one GPU through the C-ABI runs on N."""

    def __init__(self, lib: VxLib, nx: int, ny: int, nz: int, rank: int, world: int, device: int = 0,
                 voxel_size: float = 0.005, tip_load: float = 1.0, material: Material = None, host_exchange: bool = False,
                 path: int = 0, overlap: bool = True, peer: bool = True, scenario: "scenarios.Scenario" = None):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = rank, world
        self.z_origin = 0
        if scenario is not None:                        # lattice extents of the scenario; the box must be completely filled
            lo3, hi3 = scenario.ijk.min(axis=0), scenario.ijk.max(axis=0)
            nx, ny, nz = (int(v) for v in (hi3 - lo3 + 1))
            if nx * ny * nz != len(scenario.ijk) or scenario.sim_id is not None:
                raise ValueError("z-slab runs need one completely filled box")
            self.z_origin, voxel_size = int(lo3[2]), scenario.voxel_size
        self.nx, self.ny, self.nz = nx, ny, nz
        self.z0, self.z1 = slab_range(nz, rank, world)
        self.lo = self.z0 - 1 if rank > 0 else self.z0             # first stored layer (ghost below)
        self.hi = self.z1 + 1 if rank < world - 1 else self.z1     # one past the last stored layer
        self.plane = nx * ny
        sim = lib.create(voxel_size, device)
        if scenario is None:
            ijk = scenarios.box_ijk(nx, ny, self.hi - self.lo, origin=(0, 0, self.lo))
            flags = np.zeros(len(ijk), np.uint32)
            flags[(ijk[:, 2] < self.z0) | (ijk[:, 2] >= self.z1)] = VF_GHOST
            sim.set_materials([material or Material(E=1e6, rho=1e3)])
            sim.set_path(path)
            sim.set_voxels(ijk, np.zeros(len(ijk), np.uint16), flags=flags)
            owned = flags == 0
            fixed = np.nonzero((ijk[:, 0] == 0) & owned)[0]
            load = np.nonzero((ijk[:, 0] == nx - 1) & owned)[0]
            ev = np.concatenate([fixed, load]).astype(np.int32)
            dof = np.concatenate([np.full(len(fixed), DOF_ALL), np.zeros(len(load))]).astype(np.uint8)
            f = np.zeros((len(ev), 3), np.float32)
            f[len(fixed):, 2] = np.float32(-tip_load / (ny * nz))
            sim.set_externals(ev, dof, f)
        else:
            sc = scenario
            z = sc.ijk[:, 2] - self.z_origin
            keep = np.nonzero((z >= self.lo) & (z < self.hi))[0]                   # caller indices of the scenario kept here ...
            order = np.lexsort((sc.ijk[keep, 0], sc.ijk[keep, 1], sc.ijk[keep, 2]))     # ... stored plane by plane, x fastest
            keep = keep[order]
            ijk = sc.ijk[keep]
            flags = np.zeros(len(ijk), np.uint32)
            zz = ijk[:, 2] - self.z_origin
            flags[(zz < self.z0) | (zz >= self.z1)] = VF_GHOST
            sim.set_materials(sc.materials)
            sim.set_path(path)
            sim.set_gravity(sc.gravity)
            sim.enable_floor(sc.floor)
            sim.set_voxels(ijk, sc.mat[keep], flags=flags)
            if len(sc.ext_voxel):                       # externals of the voxels this rank OWNS (a ghost's pose comes from its owner)
                local = np.full(len(sc.ijk), -1, np.int64); local[keep] = np.arange(len(keep))
                sel = np.nonzero((local[sc.ext_voxel] >= 0) & (flags[np.maximum(local[sc.ext_voxel], 0)] == 0))[0]
                pick = lambda a: None if a is None else np.asarray(a)[sel]
                if len(sel):
                    sim.set_externals(local[sc.ext_voxel][sel].astype(np.int32), sc.ext_dof[sel], pick(sc.ext_force), pick(sc.ext_moment),
                                      pick(sc.ext_translation), pick(sc.ext_rotation))
            if sc.collisions:
                raise ValueError("self-collisions across z-slabs are not supported (SURVEY 8e: not required by C5b)")
            if sc.temperature is not None:
                sim.set_temperature_all(sc.temperature)
            self.scenario_index = keep                  # local voxel -> caller index of the scenario
        self.sim, self.ijk = sim, ijk
        # Poisson materials: a ghost's Poisson strain (it lacks the links to compute it) travels with the halo -- in the peer
        # stores of the step kernel, or as a third field of the host exchange; the NCCL pose messages do not carry it
        self.poisson = any(m.nu != 0.0 for m in (scenario.materials if scenario is not None else [material or Material()]))
        self.host_exchange = host_exchange
        self._bufs = None
        # overlap the exchange with the interior of the step (device exchange on the fused lattice path only)
        self.overlap = overlap and not host_exchange and world > 1 and path in (0, 5, 7) and sim.active_path() == 2
        self._comm = None
        # peer-memory halo: boundary poses stored straight into the neighbours' ghost layers (CUDA IPC over NVLink),
        # stepping loop in the library (vx_slab_step); falls back to NCCL send/recv when the mappings cannot be made
        self.peer = False
        if peer and self.overlap and dist.is_available() and dist.is_initialized():
            self.peer = self._connect_peers()

    @classmethod
    def from_scenario(cls, lib: VxLib, sc: "scenarios.Scenario", rank: int, world: int, **kw) -> "SlabRunner":
        return cls(lib, 0, 0, 0, rank, world, scenario=sc, **kw)

    def set_temperature_all(self, t: float):
        """CVoxelyze::setAmbientTemperature on every slab (ghost temperatures travel with the halo)."""
        self.sim.set_temperature_all(t)

    def _reduce_divergence(self, div):
        """doTimeStep returns false on every rank as soon as one slab diverged (src/Voxelyze.cpp:265-269): the step index is
        MIN-reduced after the call.  (The other slabs may have advanced past that step; a diverged simulation is dead anyway.)"""
        if self.world == 1 or not (self.dist.is_available() and self.dist.is_initialized()):
            return div
        import torch
        dev = "cpu" if self.host_exchange else "cuda"
        t = torch.tensor([2 ** 31 - 1 if div is None else int(div)], dtype=torch.int64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        v = int(t.item())
        return None if v == 2 ** 31 - 1 else v

    # ---- bookkeeping -------------------------------------------------------------------
    def global_counts(self):
        nx, ny, nz = self.nx, self.ny, self.nz
        return nx * ny * nz, (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1)

    def local_counts(self):
        return self.plane * (self.z1 - self.z0), self.sim.n_links

    def layer_index_range(self, z: int) -> Tuple[int, int]:
        """Voxel index range (caller order) of stored layer z."""
        first = (z - self.lo) * self.plane
        return first, self.plane

    def recommended_dt(self) -> float:
        dt = self.sim.recommended_dt()
        if self.world > 1:
            import torch
            dev = "cpu" if self.host_exchange else "cuda"
            t = torch.tensor([dt], dtype=torch.float32, device=dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
            dt = float(t.item())
        return dt

    # ---- peer-memory halo ------------------------------------------------------------
    def _connect_peers(self) -> bool:
        import torch
        from .capi import VxError
        dist, sim, n = self.dist, self.sim, Sim.PEER_DESC_BYTES
        mine = torch.zeros(2, n, dtype=torch.uint8)
        if self.rank > 0:
            mine[0] = torch.frombuffer(bytearray(sim.peer_export(self.z_origin + self.z0 - 1, False)), dtype=torch.uint8)
        if self.rank < self.world - 1:
            mine[1] = torch.frombuffer(bytearray(sim.peer_export(self.z_origin + self.z1, True)), dtype=torch.uint8)
        everyone = [torch.empty(2, n, dtype=torch.uint8, device="cuda") for _ in range(self.world)]
        dist.all_gather(everyone, mine.cuda())
        ok = 1
        try:
            if self.rank > 0:          # my first layer is the top ghost of the slab below
                sim.peer_attach(self.z_origin + self.z0, bytes(everyone[self.rank - 1][1].cpu().numpy()))
            if self.rank < self.world - 1:
                sim.peer_attach(self.z_origin + self.z1 - 1, bytes(everyone[self.rank + 1][0].cpu().numpy()))
        except VxError as e:
            self.peer_error = str(e)
            ok = 0
        t = torch.tensor([ok], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) == 0:
            sim.peer_detach()
            return False
        return True

    @staticmethod
    def connect_local(runners):
        """Wires the slabs of ONE process to each other (tests on a single GPU)."""
        for lo, hi in zip(runners[:-1], runners[1:]):
            hi.sim.peer_attach(hi.z_origin + hi.z0, lo.sim.peer_export(lo.z_origin + lo.z1, True))
            lo.sim.peer_attach(lo.z_origin + lo.z1 - 1, hi.sim.peer_export(hi.z_origin + hi.z0 - 1, False))
        for r in runners:
            r.peer = True

    # ---- halo exchange -----------------------------------------------------------------
    def _neighbours(self):
        """(peer rank, layer I send, ghost layer I receive into)"""
        out = []
        if self.rank > 0:
            out.append((self.rank - 1, self.z0, self.z0 - 1))
        if self.rank < self.world - 1:
            out.append((self.rank + 1, self.z1 - 1, self.z1))
        return out

    def _exchange_device(self, import_stream: Optional[int] = None):
        """Queues send/recv of the boundary pose planes and the ghost imports after the work already queued on
        torch's current stream; import_stream: cudaStream_t the import kernels go to (None: the handle's)."""
        import torch
        dist = self.dist
        if self.poisson:
            raise RuntimeError("Poisson materials on z-slabs need the peer-memory halo or the host exchange (the NCCL pose messages carry no Poisson strains)")
        if self._bufs is None:
            self._bufs = {"views": {}, "recv": {}}
        views, recv = self._bufs["views"], self._bufs["recv"]
        ops, imports = [], []
        for peer, send_z, recv_z in self._neighbours():
            # the fused lattice path ping-pongs generations: the current-state arrays move every step
            p0, p1, n, rb = self.sim.pose_plane(self.z_origin + send_z)
            assert n == self.plane
            for ptr in (p0, p1):
                if ptr not in views:
                    views[ptr] = torch.as_tensor(_DevMem(ptr, n * rb), device="cuda")
            if peer not in recv:
                recv[peer] = (torch.empty(n * rb, dtype=torch.uint8, device="cuda"),
                              torch.empty(n * rb, dtype=torch.uint8, device="cuda"))
            r0, r1 = recv[peer]
            ops += [dist.P2POp(dist.isend, views[p0], peer), dist.P2POp(dist.isend, views[p1], peer),
                    dist.P2POp(dist.irecv, r0, peer), dist.P2POp(dist.irecv, r1, peer)]
            imports.append((recv_z, r0, r1))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for recv_z, r0, r1 in imports:
            self.sim.halo_import(self.z_origin + recv_z, r0.data_ptr(), r1.data_ptr(), self.plane, import_stream)

    def _exchange_host(self):
        import torch
        dist = self.dist
        reqs, recvs = [], []
        for peer, send_z, recv_z in self._neighbours():
            first, n = self.layer_index_range(send_z)
            parts = [self.sim.download("pos", first, n).ravel(), self.sim.download("orient", first, n).ravel()]
            if self.poisson:
                parts.append(self.sim.download("pstrain", first, n).ravel().astype(np.float64))      # float32 values, exact in float64
            payload = np.concatenate(parts)
            reqs.append(dist.isend(torch.from_numpy(payload), peer))
            buf = torch.empty((10 if self.poisson else 7) * self.plane, dtype=torch.float64)
            reqs.append(dist.irecv(buf, peer))
            recvs.append((buf, recv_z))
        for r in reqs:
            r.wait()
        for buf, recv_z in recvs:
            first, n = self.layer_index_range(recv_z)
            a = buf.numpy()
            self.sim.upload("pos", a[:3 * n].reshape(n, 3), first)
            self.sim.upload("orient", a[3 * n:7 * n].reshape(n, 4), first)
            if self.poisson:
                self.sim.upload("pstrain", a[7 * n:].astype(np.float32).reshape(n, 3), first)

    def exchange(self):
        if self.world == 1:
            return
        if self.peer:
            self.sim.slab_exchange()
        elif self.host_exchange:
            self._exchange_host()
        else:
            self._exchange_device()

    # ---- stepping ------------------------------------------------------------------------
    def step(self, dt: float, n: int, check_divergence: bool = False):
        """n steps with their halo exchanges.  check_divergence: also agree on the divergence flag across ranks (one tiny
        all-reduce after the call; bench.py's timed region leaves it off like it leaves every other collective off)."""
        if self.peer:
            div = self.sim.slab_step(dt, n)
        elif self.overlap:
            div = self._step_overlapped(dt, n)
        else:
            div = None
            for k in range(n):
                d = self.sim.step(dt, 1)
                if d is not None and div is None:
                    div = k
                self.exchange()
        return self._reduce_divergence(div) if check_divergence else div

    def _step_overlapped(self, dt: float, n: int):
        """boundary layers first, halo push on a second stream, interior meanwhile.
        The host never blocks inside the loop; the handle must run on torch's current stream (bench.py, tests)."""
        import torch
        main = torch.cuda.current_stream()
        if self._comm is None:
            self._comm = torch.cuda.Stream()
            self.sim.set_stream(main.cuda_stream)
        comm = self._comm
        sim = self.sim
        sim.step_begin(dt)
        imported = None
        for _ in range(n):
            if imported is not None:
                main.wait_event(imported)                 # the boundary part reads the ghost poses of the previous step
            sim.step_enqueue(sim.PART_Z_BOUNDARY)
            ready = torch.cuda.Event(); ready.record(main)
            sim.step_enqueue(sim.PART_Z_INTERIOR)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                self._exchange_device(import_stream=comm.cuda_stream)
                imported = torch.cuda.Event(); imported.record(comm)
        if imported is not None:
            main.wait_event(imported)
        return sim.step_end()

    def step_profile(self, dt: float, n: int):
        tot, launches = None, None
        for _ in range(n):
            ms, ln = self.sim.step_profile(dt, 1)
            self.exchange()                              # un-overlapped here: the kernel is what is being timed
            tot = ms if tot is None else {k: tot[k] + ms[k] for k in ms}
            launches = ln if launches is None else [a + b for a, b in zip(launches, ln)]
        return tot, launches

    def read_probe(self):
        first, n = self.layer_index_range(self.z1 - 1)
        return self.sim.download("pos", first + n - 1, 1)

    def owned_state(self, field: str) -> np.ndarray:
        first, _ = self.layer_index_range(self.z0)
        return self.sim.download(field, first, self.plane * (self.z1 - self.z0))

    def dominant_kernel(self) -> str:
        return _kernel_name(self.sim)

    def path_name(self) -> str:
        return f"{_path_name(self.sim)}, z-slab {self.rank}/{self.world} layers [{self.z0},{self.z1})"

    def h2d_bytes_per_step(self) -> int:
        return 4

    def close(self):
        if self.peer:
            self.sim.peer_detach()
        self.sim.close()

    def halo_name(self) -> str:
        if self.peer:
            return "peer stores into the neighbours' ghost layers over NVLink (CUDA IPC), overlapped with the interior part"
        if self.overlap:
            return "NCCL send/recv on a second stream, overlapped with the interior part"
        return "host copies (gloo)" if self.host_exchange else "NCCL send/recv after the step"
