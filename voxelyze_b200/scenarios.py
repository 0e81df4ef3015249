"""Synthetic lattices: the BASELINE.json configs and small test cases.

A :class:`Scenario` is the flat description a caller of the reference would build with
``addMaterial`` / ``setVoxel`` / ``external()->set*`` (README.md:20-45 of the reference);
:func:`build` replays it on any implementation of the C-ABI.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .capi import Material, Sim, VxLib, DOF_ALL, MODEL_BILINEAR


@dataclass
class Scenario:
    name: str
    voxel_size: float
    materials: List[Material]
    ijk: np.ndarray                 # (n,3) int32, setVoxel order
    mat: np.ndarray                 # (n,) uint16
    sim_id: Optional[np.ndarray] = None
    ext_voxel: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    ext_dof: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))
    ext_force: Optional[np.ndarray] = None
    ext_moment: Optional[np.ndarray] = None
    ext_translation: Optional[np.ndarray] = None
    ext_rotation: Optional[np.ndarray] = None
    gravity: float = 0.0
    floor: bool = False
    collisions: bool = False
    temperature: Optional[float] = None
    dt: Optional[float] = None      # None: use recommended dt

    @property
    def n_voxels(self) -> int:
        return len(self.ijk)


def build(lib: VxLib, sc: Scenario, device: int = 0, path: int = 0) -> Sim:
    """path: 0 auto, 1 force the general two-kernel path, 2 fused lattice path (CUDA library only)."""
    s = lib.create(sc.voxel_size, device)
    s.set_materials(sc.materials)
    s.set_path(path)
    s.set_gravity(sc.gravity)
    s.enable_floor(sc.floor)
    s.set_voxels(sc.ijk, sc.mat, sc.sim_id)
    if len(sc.ext_voxel):
        s.set_externals(sc.ext_voxel, sc.ext_dof, sc.ext_force, sc.ext_moment,
                        sc.ext_translation, sc.ext_rotation)
    if sc.collisions:
        s.enable_collisions(True)
    if sc.temperature is not None:
        s.set_temperature_all(sc.temperature)
    return s


def build_slabbed(lib: VxLib, sc: Scenario, devices) -> "SlabbedSim":
    """The same calls as build() on a vx_slabbed handle: the whole model in the caller's numbering, cut into z-slabs over
    `devices` of this process (include/voxelyze_b200.h)."""
    if sc.sim_id is not None or sc.collisions:
        raise ValueError("vx_slabbed: one body, no self-collisions")
    s = lib.create_slabbed(sc.voxel_size, devices)
    s.set_materials(sc.materials)
    s.set_gravity(sc.gravity)
    s.enable_floor(sc.floor)
    s.set_voxels(sc.ijk, sc.mat)
    if len(sc.ext_voxel):
        s.set_externals(sc.ext_voxel, sc.ext_dof, sc.ext_force, sc.ext_moment, sc.ext_translation, sc.ext_rotation)
    if sc.temperature is not None:
        s.set_temperature_all(sc.temperature)
    return s


def box_ijk(nx: int, ny: int, nz: int, origin=(0, 0, 0)) -> np.ndarray:
    """Lattice indices of a solid box in the insertion order k -> j -> i (x fastest)."""
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    out = np.stack([i.ravel() + origin[0], j.ravel() + origin[1], k.ravel() + origin[2]], axis=1)
    return out.astype(np.int32)


def _externals(sc: Scenario, fixed_idx, load_idx, load_force):
    ev = np.concatenate([fixed_idx, load_idx]).astype(np.int32)
    dof = np.concatenate([np.full(len(fixed_idx), DOF_ALL), np.zeros(len(load_idx))]).astype(np.uint8)
    f = np.zeros((len(ev), 3), np.float32)
    f[len(fixed_idx):] = np.asarray(load_force, np.float32)
    sc.ext_voxel, sc.ext_dof, sc.ext_force = ev, dof, f
    return sc


def cantilever(nx=20, ny=4, nz=4, voxel_size=0.005, E=1e6, rho=1e3, tip_load=1.0,
               name=None) -> Scenario:
    """C1 / C5 pattern: x=0 face fixed, total `tip_load` N in -z shared by the x=nx-1 face.

    C1 = cantilever() (SURVEY.md section 8c anchor); C5a = cantilever(256,256,256) with
    tip_load=1 giving -1/65536 N per face voxel."""
    ijk = box_ijk(nx, ny, nz)
    sc = Scenario(name or f"cantilever_{nx}x{ny}x{nz}", voxel_size, [Material(E=E, rho=rho)],
                  ijk, np.zeros(len(ijk), np.uint16))
    fixed = np.nonzero(ijk[:, 0] == 0)[0]
    load = np.nonzero(ijk[:, 0] == nx - 1)[0]
    return _externals(sc, fixed, load, [0.0, 0.0, -tip_load / (ny * nz)])


def drop_block(n=64, voxel_size=0.005, name=None) -> Scenario:
    """C2: n^3 block one voxel above the floor, gravity, global + collision damping."""
    ijk = box_ijk(n, n, n, origin=(0, 0, 1))
    m = Material(E=1e6, rho=1e3, zeta_global=0.01, zeta_collision=1.0)
    return Scenario(name or f"drop_{n}", voxel_size, [m], ijk, np.zeros(len(ijk), np.uint16),
                    gravity=1.0, floor=True)


def plate_stack(nx=128, ny=128, nplates=16, thick=6, gap=2, voxel_size=0.005, tip_load=None,
                checker=4, name=None) -> Scenario:
    """C3: stack of cantilever plates separated by empty layers, two bilinear materials in a
    checkerboard, gravity + floor + self collisions."""
    A = Material(model=MODEL_BILINEAR, E=1e6, plastic_modulus=2e5, yield_stress=4e3, fail_stress=8e3,
                 rho=1e3, zeta_global=0.01, zeta_collision=0.5)
    B = Material(model=MODEL_BILINEAR, E=1e7, plastic_modulus=2e6, yield_stress=4e4, fail_stress=8e4,
                 rho=1e3, zeta_global=0.01, zeta_collision=0.5)
    period = thick + gap
    parts = [box_ijk(nx, ny, thick, origin=(0, 0, p * period)) for p in range(nplates)]
    ijk = np.concatenate(parts)
    mat = (((ijk[:, 0] // checker) + (ijk[:, 1] // checker) + (ijk[:, 2] // checker)) & 1).astype(np.uint16)
    sc = Scenario(name or f"plates_{nx}x{ny}x{nplates}", voxel_size, [A, B], ijk, mat,
                  gravity=1.0, floor=True, collisions=True)
    plate = ijk[:, 2] // period
    fixed = np.nonzero((ijk[:, 0] == 0) & (plate >= 1))[0]
    load = np.nonzero((ijk[:, 0] == nx - 1) & (plate == nplates - 1))[0]
    if tip_load is None:
        tip_load = 1.0 * (ny / 8.0)
    return _externals(sc, fixed, load, [0.0, 0.0, -tip_load / max(1, len(load))])


def robot_materials() -> List[Material]:
    kw = dict(rho=1e3, zeta_global=0.05, mu_static=1.0, mu_kinetic=0.5)
    return [Material(E=1e6, cte=0.01, **kw), Material(E=1e6, cte=-0.01, **kw), Material(E=5e6, cte=0.0, **kw)]


def robot_ensemble(n_robots=4096, n=10, voxel_size=0.005, first_seed=0, name=None) -> Scenario:
    """C4: independent n^3 robots, material by hash of (i,j,k,seed), CTE actuation."""
    base = box_ijk(n, n, n)
    ijk = np.tile(base, (n_robots, 1))
    seeds = np.repeat(np.arange(first_seed, first_seed + n_robots), len(base))
    h = (ijk[:, 0] * 7 + ijk[:, 1] * 13 + ijk[:, 2] * 29 + seeds * 31) % 3
    return Scenario(name or f"robots_{n_robots}x{n}", voxel_size, robot_materials(), ijk,
                    h.astype(np.uint16), sim_id=(seeds - first_seed).astype(np.int32),
                    gravity=1.0, floor=True)


def robot_temperature(t: float) -> float:
    """Ambient temperature program of C4: 20*sinf(2*pi*40*t), float arithmetic."""
    return float(np.float32(20.0) * np.sin(np.float32(2.0 * np.pi * 40.0) * np.float32(t), dtype=np.float32))
