"""ctypes binding of include/voxelyze_b200.h.

The same binding drives three shared objects that export the identical C-ABI:

* ``voxelyze_b200/lib/libvoxelyze_b200.so`` -- the product (sm_100a CUDA kernels)
* ``oracle/liboracle_port.so``              -- CPU restatement (tests/bench baseline only)
* ``oracle/_ref/libvxref.so``               -- unmodified reference behind a shim (ditto)

Only :func:`load_product` is product code; :func:`load_oracle` / :func:`load_reference`
exist for tests, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline legs.
There is no fallback between them: a missing CUDA library raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_SO = os.path.join(ROOT, "voxelyze_b200", "lib", "libvoxelyze_b200.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle_port.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libvxref.so")
REF_OMP_SO = os.path.join(ROOT, "oracle", "_ref", "libvxref_omp.so")

VX_OK, VX_DIVERGED = 0, 1
MODEL_LINEAR, MODEL_BILINEAR, MODEL_DATA = 0, 1, 2

# enum vx_field -> (numpy dtype, components, is_link)
FIELDS = {
    "pos": (0, np.float64, 3, False), "orient": (1, np.float64, 4, False),
    "linmom": (2, np.float64, 3, False), "angmom": (3, np.float64, 3, False),
    "temp": (4, np.float32, 1, False), "voxflags": (5, np.uint32, 1, False),
    "pstrain": (6, np.float32, 3, False),
    "force_neg": (16, np.float64, 3, True), "force_pos": (17, np.float64, 3, True),
    "moment_neg": (18, np.float64, 3, True), "moment_pos": (19, np.float64, 3, True),
    "pos2": (20, np.float64, 3, True), "angle1v": (21, np.float64, 3, True),
    "angle2v": (22, np.float64, 3, True),
    "strain": (24, np.float32, 1, True), "maxstrain": (25, np.float32, 1, True),
    "strainoffset": (26, np.float32, 1, True), "stress": (27, np.float32, 1, True),
    "linkflags": (28, np.uint32, 1, True),
}
VF_STATIC_FRICTION, VF_SURFACE, VF_GHOST = 1, 2, 4
LF_SMALL_ANGLE, LF_LOCAL_VEL_VALID, LF_YIELDED, LF_FAILED = 1, 2, 4, 8
DOF_ALL = 0x3F


class MaterialDesc(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("youngs_modulus", C.c_float), ("plastic_modulus", C.c_float),
        ("yield_stress", C.c_float), ("fail_stress", C.c_float), ("n_points", C.c_int32),
        ("strain", C.POINTER(C.c_float)), ("stress", C.POINTER(C.c_float)),
        ("density", C.c_float), ("poissons_ratio", C.c_float), ("cte", C.c_float),
        ("mu_static", C.c_float), ("mu_kinetic", C.c_float), ("zeta_internal", C.c_float),
        ("zeta_global", C.c_float), ("zeta_collision", C.c_float), ("ext_scale", C.c_double * 3),
    ]


class VoxMatRow(C.Structure):
    _fields_ = [("nom_size", C.c_double), ("size", C.c_double * 3)] + [
        (n, C.c_float) for n in (
            "E", "nu", "rho", "cte", "mu_static", "mu_kinetic", "zeta_internal", "zeta_global",
            "zeta_collision", "e_hat", "mass", "mass_inv", "sqrt_mass", "first_moment",
            "moment_inertia", "moment_inertia_inv", "two_sq_m_e_s", "two_sq_i_e_s3",
            "eps_yield", "eps_fail", "sigma_yield", "sigma_fail")
    ] + [("linear", C.c_int32), ("n_curve", C.c_int32)]


class LinkMatRow(C.Structure):
    _fields_ = [("mat_a", C.c_int32), ("mat_b", C.c_int32), ("linear", C.c_int32), ("n_curve", C.c_int32)] + [
        (n, C.c_float) for n in (
            "E", "nu", "e_hat", "eps_yield", "eps_fail", "sigma_yield", "sigma_fail",
            "a1", "a2", "b1", "b2", "b3", "sq_a1", "sq_a2_ip", "sq_b1", "sq_b2_fmp", "sq_b3_ip")
    ]


def row_to_dict(row: C.Structure) -> dict:
    out = {}
    for name, _ in row._fields_:
        v = getattr(row, name)
        out[name] = list(v) if hasattr(v, "__len__") else v
    return out


@dataclass
class Material:
    """User-level material, mirrors the setters of CVX_Material (include/VX_Material.h:33-103)."""
    E: float = 1e6
    rho: float = 1e3
    model: int = MODEL_LINEAR
    plastic_modulus: float = 0.0
    yield_stress: float = 0.0
    fail_stress: float = -1.0
    strain: Sequence[float] = ()
    stress: Sequence[float] = ()
    nu: float = 0.0
    cte: float = 0.0
    mu_static: float = 0.0
    mu_kinetic: float = 0.0
    zeta_internal: float = 1.0
    zeta_global: float = 0.0
    zeta_collision: float = 0.0
    ext_scale: Sequence[float] = (1.0, 1.0, 1.0)


class VxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vx error {code}: {msg}")
        self.code = code


class VxLib:
    """One loaded implementation of the C-ABI."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing - build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L = self.lib
        vp, i32, f32, u64 = C.c_void_p, C.c_int, C.c_float, C.c_uint64
        P = C.POINTER
        sig = {
            "vx_create": (i32, [C.c_double, i32, P(vp)]),
            "vx_destroy": (None, [vp]),
            "vx_last_error": (C.c_char_p, [vp]),
            "vx_backend": (C.c_char_p, []),
            "vx_abi_version": (i32, []),
            "vx_set_materials": (i32, [vp, i32, P(MaterialDesc)]),
            "vx_get_voxmat": (i32, [vp, i32, P(VoxMatRow)]),
            "vx_get_linkmat": (i32, [vp, i32, i32, P(LinkMatRow)]),
            "vx_get_linkmat_curve": (i32, [vp, i32, i32, vp, vp, i32]),
            "vx_set_voxels": (i32, [vp, i32, vp, vp, vp, vp]),
            "vx_voxel_count": (i32, [vp]),
            "vx_link_count": (i32, [vp]),
            "vx_get_links": (i32, [vp, vp, vp, vp]),
            "vx_set_externals": (i32, [vp, i32, vp, vp, vp, vp, vp, vp]),
            "vx_set_gravity": (i32, [vp, f32]),
            "vx_enable_floor": (i32, [vp, i32]),
            "vx_enable_collisions": (i32, [vp, i32]),
            "vx_set_collision_envelope": (i32, [vp, f32]),
            "vx_set_temperature_all": (i32, [vp, f32]),
            "vx_set_temperature_members": (i32, [vp, i32, vp]),
            "vx_set_temperature": (i32, [vp, i32, vp]),
            "vx_step": (i32, [vp, f32, i32, P(i32)]),
            "vx_step_ambient": (i32, [vp, f32, i32, vp, P(i32)]),
            "vx_prepare": (i32, [vp]),
            "vx_recommended_dt": (i32, [vp, P(f32)]),
            "vx_reset": (i32, [vp]),
            "vx_time": (f32, [vp]),
            "vx_set_clock": (i32, [vp, f32, f32]),
            "vx_download": (i32, [vp, i32, i32, i32, vp]),
            "vx_upload": (i32, [vp, i32, i32, i32, vp]),
            "vx_download_voxel_state": (i32, [vp, i32, i32, vp]),
            "vx_collision_pairs": (i32, [vp, vp, i32, P(i32)]),
            "vx_collision_stats": (i32, [vp, P(i32), P(i32)]),
            "vx_state_info": (i32, [vp, i32, i32, P(f32)]),
            "vx_linear_solve": (i32, [vp, C.c_double, i32, P(i32), P(C.c_double)]),
            "vx_mesh_set_material_colors": (i32, [vp, i32, vp]),
            "vx_mesh_build": (i32, [vp, P(i32), P(i32)]),
            "vx_mesh_update": (i32, [vp, i32, i32]),
            "vx_mesh_counts": (i32, [vp, P(i32), P(i32)]),
            "vx_mesh_download": (i32, [vp, vp, vp, vp, vp, vp]),
            "vx_mesh_device": (i32, [vp, P(u64), P(u64), P(u64), P(u64)]),
            "vx_set_stream": (i32, [vp, u64]),
            "vx_pose_plane": (i32, [vp, i32, P(u64), P(u64), P(i32), P(i32)]),
            "vx_halo_import": (i32, [vp, i32, u64, u64, i32]),
            "vx_launch_count": (C.c_int64, [vp]),
            "vx_sync": (i32, [vp]),
            "vx_halo_import_on": (i32, [vp, i32, C.c_uint64, C.c_uint64, i32, C.c_uint64]),
            "vx_step_begin": (i32, [vp, C.c_float]),
            "vx_step_enqueue": (i32, [vp, i32]),
            "vx_step_end": (i32, [vp, C.POINTER(i32)]),
            "vx_peer_export": (i32, [vp, i32, i32, C.c_void_p]),
            "vx_peer_attach": (i32, [vp, i32, C.c_void_p]),
            "vx_peer_detach": (i32, [vp]),
            "vx_slab_step": (i32, [vp, C.c_float, i32, C.POINTER(i32)]),
            "vx_slab_exchange": (i32, [vp]),
            "vx_save_state": (i32, [vp, C.c_char_p]),
            "vx_load_state": (i32, [vp, C.c_char_p]),
            "vx_download_link_state": (i32, [vp, i32, i32, C.c_void_p]),
            "vx_upload_link_state": (i32, [vp, i32, i32, C.c_void_p]),
            "vx_collision_forces": (i32, [vp, C.c_void_p, C.c_void_p, i32, C.POINTER(i32)]),
            "vx_set_path": (i32, [vp, i32]),
            "vx_active_path": (i32, [vp]),
            "vx_kernel_name": (C.c_char_p, [vp]),
            "vx_step_profile": (i32, [vp, f32, i32, P(f32), P(i32)]),
            "vx_slab_step_begin": (i32, [vp, f32, i32]),
            "vx_slab_step_finish": (i32, [vp, P(i32)]),
            # one lattice on several devices of one process
            "vx_slabbed_create": (i32, [C.c_double, i32, vp, P(vp)]),
            "vx_slabbed_destroy": (None, [vp]),
            "vx_slabbed_last_error": (C.c_char_p, [vp]),
            "vx_slabbed_slab_count": (i32, [vp]),
            "vx_slabbed_slab": (vp, [vp, i32]),
            "vx_slabbed_halo_mode": (i32, [vp]),
            "vx_slabbed_set_materials": (i32, [vp, i32, P(MaterialDesc)]),
            "vx_slabbed_set_gravity": (i32, [vp, f32]),
            "vx_slabbed_enable_floor": (i32, [vp, i32]),
            "vx_slabbed_set_voxels": (i32, [vp, i32, vp, vp]),
            "vx_slabbed_voxel_count": (i32, [vp]),
            "vx_slabbed_link_count": (i32, [vp]),
            "vx_slabbed_get_links": (i32, [vp, vp, vp, vp]),
            "vx_slabbed_set_externals": (i32, [vp, i32, vp, vp, vp, vp, vp, vp]),
            "vx_slabbed_set_temperature_all": (i32, [vp, f32]),
            "vx_slabbed_set_temperature": (i32, [vp, i32, vp]),
            "vx_slabbed_step": (i32, [vp, f32, i32, P(i32)]),
            "vx_slabbed_recommended_dt": (i32, [vp, P(f32)]),
            "vx_slabbed_reset": (i32, [vp]),
            "vx_slabbed_time": (f32, [vp]),
            "vx_slabbed_set_clock": (i32, [vp, f32, f32]),
            "vx_slabbed_download": (i32, [vp, i32, i32, i32, vp]),
            "vx_slabbed_upload": (i32, [vp, i32, i32, i32, vp]),
            "vx_slabbed_download_voxel_state": (i32, [vp, i32, i32, vp]),
            "vx_slabbed_download_link_state": (i32, [vp, i32, i32, vp]),
            "vx_slabbed_upload_link_state": (i32, [vp, i32, i32, vp]),
            "vx_slabbed_state_info": (i32, [vp, i32, i32, P(f32)]),
            "vx_slabbed_save_state": (i32, [vp, C.c_char_p]),
            "vx_slabbed_load_state": (i32, [vp, C.c_char_p]),
            "vx_slabbed_launch_count": (C.c_int64, [vp]),
        }
        self.symbols = list(sig)
        for name, (res, args) in sig.items():
            fn = getattr(L, name)     # AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args

    @property
    def backend(self) -> str:
        return self.lib.vx_backend().decode()

    def create(self, voxel_size: float, device: int = 0) -> "Sim":
        return Sim(self, voxel_size, device)

    def create_slabbed(self, voxel_size: float, devices: Sequence[int]) -> "SlabbedSim":
        return SlabbedSim(self, voxel_size, devices)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Sim:
    """Thin object wrapper over one vx_sim handle."""

    def __init__(self, lib: VxLib, voxel_size: float, device: int = 0):
        self.L = lib
        self.h = C.c_void_p()
        rc = lib.lib.vx_create(float(voxel_size), int(device), C.byref(self.h))
        if rc != VX_OK:
            raise VxError(rc, f"vx_create failed on {lib.path}")
        self.voxel_size = voxel_size
        self._keep = []

    def close(self):
        if self.h:
            self.L.lib.vx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc: int, ok=(VX_OK,)):
        if rc not in ok:
            raise VxError(rc, self.L.lib.vx_last_error(self.h).decode())
        return rc

    # -- model -----------------------------------------------------------------
    def set_materials(self, mats: Sequence[Material]):
        arr, keep = _material_descs(mats)
        self._chk(self.L.lib.vx_set_materials(self.h, len(mats), arr))

    def voxmat(self, i: int) -> dict:
        r = VoxMatRow()
        self._chk(self.L.lib.vx_get_voxmat(self.h, i, C.byref(r)))
        return row_to_dict(r)

    def linkmat(self, a: int, b: int) -> dict:
        r = LinkMatRow()
        self._chk(self.L.lib.vx_get_linkmat(self.h, a, b, C.byref(r)))
        d = row_to_dict(r)
        n = d["n_curve"]
        s = np.zeros(n, np.float32)
        t = np.zeros(n, np.float32)
        got = self.L.lib.vx_get_linkmat_curve(self.h, a, b, _ptr(s), _ptr(t), n)
        if got < 0:
            self._chk(got)
        d["curve_strain"], d["curve_stress"] = s, t
        return d

    def set_voxels(self, ijk, mat, sim_id=None, flags=None):
        ijk = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        mat = np.ascontiguousarray(mat, dtype=np.uint16)
        sid = None if sim_id is None else np.ascontiguousarray(sim_id, dtype=np.int32)
        fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint32)
        self._chk(self.L.lib.vx_set_voxels(self.h, len(ijk), _ptr(ijk), _ptr(mat), _ptr(sid), _ptr(fl)))

    @property
    def n_voxels(self) -> int:
        return self.L.lib.vx_voxel_count(self.h)

    @property
    def n_links(self) -> int:
        return self.L.lib.vx_link_count(self.h)

    def links(self):
        n = self.n_links
        vn, vp, ax = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
        self._chk(self.L.lib.vx_get_links(self.h, _ptr(vn), _ptr(vp), _ptr(ax)))
        return vn, vp, ax

    def set_externals(self, voxel, dof, force=None, moment=None, translation=None, rotation=None):
        voxel = np.ascontiguousarray(voxel, dtype=np.int32)
        n = len(voxel)
        dof = np.ascontiguousarray(dof, dtype=np.uint8)
        f = None if force is None else np.ascontiguousarray(force, dtype=np.float32).reshape(n, 3)
        m = None if moment is None else np.ascontiguousarray(moment, dtype=np.float32).reshape(n, 3)
        t = None if translation is None else np.ascontiguousarray(translation, dtype=np.float64).reshape(n, 3)
        r = None if rotation is None else np.ascontiguousarray(rotation, dtype=np.float64).reshape(n, 3)
        self._chk(self.L.lib.vx_set_externals(self.h, n, _ptr(voxel), _ptr(dof), _ptr(f), _ptr(m), _ptr(t), _ptr(r)))

    def set_gravity(self, g: float):
        self._chk(self.L.lib.vx_set_gravity(self.h, g))

    def enable_floor(self, on: bool = True):
        self._chk(self.L.lib.vx_enable_floor(self.h, int(on)))

    def enable_collisions(self, on: bool = True):
        self._chk(self.L.lib.vx_enable_collisions(self.h, int(on)))

    def set_temperature_all(self, t: float):
        self._chk(self.L.lib.vx_set_temperature_all(self.h, t))

    def set_temperature_members(self, t):
        t = np.ascontiguousarray(t, dtype=np.float32)
        self._chk(self.L.lib.vx_set_temperature_members(self.h, len(t), _ptr(t)))

    def set_temperature(self, t):
        t = np.ascontiguousarray(t, dtype=np.float32)
        self._chk(self.L.lib.vx_set_temperature(self.h, len(t), _ptr(t)))

    # -- hot path --------------------------------------------------------------
    def step(self, dt: float, n: int = 1) -> Optional[int]:
        """Runs n steps; returns None, or the number of completed steps if diverged."""
        div = C.c_int(-1)
        rc = self._chk(self.L.lib.vx_step(self.h, dt, n, C.byref(div)), ok=(VX_OK, VX_DIVERGED))
        return div.value if rc == VX_DIVERGED else None

    def step_ambient(self, dt: float, ambient) -> Optional[int]:
        """len(ambient) steps; before step k every voxel takes temperature ambient[k] (setAmbientTemperature + doTimeStep in turn)."""
        a = np.ascontiguousarray(ambient, dtype=np.float32)
        div = C.c_int(-1)
        rc = self._chk(self.L.lib.vx_step_ambient(self.h, dt, len(a), _ptr(a), C.byref(div)), ok=(VX_OK, VX_DIVERGED))
        return div.value if rc == VX_DIVERGED else None

    def prepare(self):
        """Builds the captured step graphs now instead of inside the first long step call."""
        self._chk(self.L.lib.vx_prepare(self.h))

    def recommended_dt(self) -> float:
        dt = C.c_float()
        self._chk(self.L.lib.vx_recommended_dt(self.h, C.byref(dt)))
        return dt.value

    def reset(self):
        self._chk(self.L.lib.vx_reset(self.h))

    def time(self) -> float:
        return self.L.lib.vx_time(self.h)

    def set_clock(self, time: float, previous_dt: float):
        self._chk(self.L.lib.vx_set_clock(self.h, time, previous_dt))

    # -- state -----------------------------------------------------------------
    def download(self, name: str, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        fid, dt, comps, is_link = FIELDS[name]
        total = self.n_links if is_link else self.n_voxels
        if count is None:
            count = total - first
        out = np.zeros((count, comps), dtype=dt)
        if count:
            self._chk(self.L.lib.vx_download(self.h, fid, first, count, _ptr(out)))
        return out if comps > 1 else out.reshape(-1)

    VOXEL_STATE_DTYPE = np.dtype([("pos", "<f8", 3), ("orient", "<f8", 4), ("linmom", "<f8", 3), ("angmom", "<f8", 3), ("temp", "<f4"), ("flags", "<u4")])

    def download_voxel_state(self, first: int = 0, count: int = 1) -> np.ndarray:
        out = np.zeros(count, self.VOXEL_STATE_DTYPE)
        self._chk(self.L.lib.vx_download_voxel_state(self.h, first, count, out.ctypes.data))
        return out

    def upload(self, name: str, data, first: int = 0):
        fid, dt, comps, _ = FIELDS[name]
        a = np.ascontiguousarray(data, dtype=dt).reshape(-1, comps)
        self._chk(self.L.lib.vx_upload(self.h, fid, first, len(a), _ptr(a)))

    def collision_pairs(self) -> np.ndarray:
        n = C.c_int(0)
        self._chk(self.L.lib.vx_collision_pairs(self.h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 2), np.int32)
        if n.value:
            self._chk(self.L.lib.vx_collision_pairs(self.h, _ptr(out), n.value, C.byref(n)))
        return out

    def collision_stats(self):
        """(watched pairs, watch-list rebuilds so far; -1 where not counted)."""
        a, b = C.c_int(0), C.c_int(0)
        self._chk(self.L.lib.vx_collision_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def linear_solve(self, rel_tol: float = 0.0, max_iter: int = 0):
        """CVoxelyze::doLinearSolve (src/Voxelyze.cpp:243-249): static solution about the nominal lattice, written into the
        poses.  Returns (iterations, relative residual); raises VxError(VX_ERR_SOLVER) when the model is not held."""
        it, res = C.c_int(0), C.c_double(0.0)
        self._chk(self.L.lib.vx_linear_solve(self.h, rel_tol, max_iter, C.byref(it), C.byref(res)))
        return it.value, res.value

    # -- surface mesh (CVX_MeshRender) --------------------------------------------
    MESH_MATERIAL, MESH_FAILURE, MESH_STATE_INFO = 0, 1, 2

    def mesh_set_material_colors(self, rgba):
        a = np.ascontiguousarray(rgba, dtype=np.uint8).reshape(-1, 4)
        self._chk(self.L.lib.vx_mesh_set_material_colors(self.h, len(a), _ptr(a)))

    def mesh(self, coloring: int = 0, state_type: int = 0) -> dict:
        """generateMesh (first call) + updateMesh; returns vertices (nv,3), quads (nq,4), normals, colors (nq,3), quad_voxel."""
        nv, nq = C.c_int(0), C.c_int(0)
        self._chk(self.L.lib.vx_mesh_counts(self.h, C.byref(nv), C.byref(nq)))
        if nv.value == 0:
            self._chk(self.L.lib.vx_mesh_build(self.h, C.byref(nv), C.byref(nq)))
        self._chk(self.L.lib.vx_mesh_update(self.h, coloring, state_type))
        out = dict(vertices=np.zeros((nv.value, 3), np.float32), quads=np.zeros((nq.value, 4), np.int32), normals=np.zeros((nq.value, 3), np.float32),
                   colors=np.zeros((nq.value, 3), np.float32), quad_voxel=np.zeros(nq.value, np.int32))
        self._chk(self.L.lib.vx_mesh_download(self.h, _ptr(out["vertices"]), _ptr(out["quads"]), _ptr(out["normals"]), _ptr(out["colors"]), _ptr(out["quad_voxel"])))
        return out

    def state_info(self, info: int, typ: int) -> float:
        v = C.c_float()
        self._chk(self.L.lib.vx_state_info(self.h, info, typ, C.byref(v)))
        return v.value

    # -- device hooks ------------------------------------------------------------
    def set_stream(self, stream):
        """stream: a cudaStream_t as int (0 = CUDA's legacy default stream), or None for the library's own stream."""
        self._chk(self.L.lib.vx_set_stream(self.h, 0xFFFFFFFFFFFFFFFF if stream is None else int(stream)))

    def pose_plane(self, iz: int):
        p0, p1, n, rb = C.c_uint64(), C.c_uint64(), C.c_int(), C.c_int()
        self._chk(self.L.lib.vx_pose_plane(self.h, iz, C.byref(p0), C.byref(p1), C.byref(n), C.byref(rb)))
        return p0.value, p1.value, n.value, rb.value

    def halo_import(self, iz: int, ptr0: int, ptr1: int, count: int, stream: Optional[int] = None):
        if stream is None:
            self._chk(self.L.lib.vx_halo_import(self.h, iz, ptr0, ptr1, count))
        else:
            self._chk(self.L.lib.vx_halo_import_on(self.h, iz, ptr0, ptr1, count, stream))

    # asynchronous call: step_begin, step_enqueue(part)..., step_end (include/voxelyze_b200.h)
    PART_ALL, PART_Z_BOUNDARY, PART_Z_INTERIOR = 0, 1, 2

    def step_begin(self, dt: float):
        self._chk(self.L.lib.vx_step_begin(self.h, dt))

    def step_enqueue(self, part: int = 0):
        self._chk(self.L.lib.vx_step_enqueue(self.h, part))

    PEER_DESC_BYTES = 512

    def peer_export(self, ghost_iz: int, from_above: bool) -> bytes:
        buf = C.create_string_buffer(self.PEER_DESC_BYTES)
        self._chk(self.L.lib.vx_peer_export(self.h, ghost_iz, 1 if from_above else 0, buf))
        return buf.raw

    def peer_attach(self, send_iz: int, desc: bytes):
        self._chk(self.L.lib.vx_peer_attach(self.h, send_iz, C.create_string_buffer(desc, self.PEER_DESC_BYTES)))

    def peer_detach(self):
        self._chk(self.L.lib.vx_peer_detach(self.h))

    def slab_step(self, dt: float, n: int = 1) -> Optional[int]:
        div = C.c_int(-1)
        rc = self._chk(self.L.lib.vx_slab_step(self.h, dt, n, C.byref(div)), ok=(VX_OK, VX_DIVERGED))
        return div.value if rc == VX_DIVERGED else None

    def slab_exchange(self):
        self._chk(self.L.lib.vx_slab_exchange(self.h))

    def step_end(self) -> Optional[int]:
        div = C.c_int(-1)
        rc = self._chk(self.L.lib.vx_step_end(self.h, C.byref(div)), ok=(VX_OK, VX_DIVERGED))
        return div.value if rc == VX_DIVERGED else None

    def launch_count(self) -> int:
        return self.L.lib.vx_launch_count(self.h)

    def sync(self):
        self._chk(self.L.lib.vx_sync(self.h))

    def step_profile(self, dt: float, n: int):
        """Event-timed steps: returns ({'link','voxel','other','step'} ms totals, launches per group)."""
        ms = (C.c_float * 4)()
        ln = (C.c_int * 3)()
        self._chk(self.L.lib.vx_step_profile(self.h, dt, n, ms, ln), ok=(VX_OK, VX_DIVERGED))
        return dict(link=ms[0], voxel=ms[1], other=ms[2], step=ms[3]), list(ln)

    def active_path(self) -> int:
        return self.L.lib.vx_active_path(self.h)

    LINK_STATE_DTYPE = np.dtype([("pos2", "<f8", 3), ("angle1v", "<f8", 3), ("angle2v", "<f8", 3), ("strain", "<f4"), ("max_strain", "<f4"),
                                 ("strain_offset", "<f4"), ("stress", "<f4"), ("flags", "<u4"), ("reserved", "<u4")])

    def download_link_state(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.n_links - first if count is None else count
        out = np.zeros(count, self.LINK_STATE_DTYPE)
        self._chk(self.L.lib.vx_download_link_state(self.h, first, count, out.ctypes.data))
        return out

    def upload_link_state(self, rec: np.ndarray, first: int = 0):
        rec = np.ascontiguousarray(rec, dtype=self.LINK_STATE_DTYPE)
        self._chk(self.L.lib.vx_upload_link_state(self.h, first, len(rec), rec.ctypes.data))

    def save_state(self, path: str):
        self._chk(self.L.lib.vx_save_state(self.h, os.fsencode(path)))

    def load_state(self, path: str):
        self._chk(self.L.lib.vx_load_state(self.h, os.fsencode(path)))

    def kernel_name(self) -> str:
        return self.L.lib.vx_kernel_name(self.h).decode()

    def set_path(self, path: int):
        self._chk(self.L.lib.vx_set_path(self.h, path))


def _material_descs(mats: Sequence[Material]):
    arr = (MaterialDesc * len(mats))()
    keep = []
    for d, m in zip(arr, mats):
        d.model = m.model
        d.youngs_modulus, d.plastic_modulus = m.E, m.plastic_modulus
        d.yield_stress, d.fail_stress = m.yield_stress, m.fail_stress
        if m.model == MODEL_DATA:
            s = np.ascontiguousarray(m.strain, dtype=np.float32)
            t = np.ascontiguousarray(m.stress, dtype=np.float32)
            keep += [s, t]
            d.n_points = len(s)
            d.strain = s.ctypes.data_as(C.POINTER(C.c_float))
            d.stress = t.ctypes.data_as(C.POINTER(C.c_float))
        d.density, d.poissons_ratio, d.cte = m.rho, m.nu, m.cte
        d.mu_static, d.mu_kinetic = m.mu_static, m.mu_kinetic
        d.zeta_internal, d.zeta_global, d.zeta_collision = m.zeta_internal, m.zeta_global, m.zeta_collision
        d.ext_scale[0], d.ext_scale[1], d.ext_scale[2] = m.ext_scale
    return arr, keep


class _SlabView(Sim):
    """One slab of a SlabbedSim, for reports (kernel_name, launch_count, active_path); owned by the slabbed handle."""

    def __init__(self, lib: VxLib, handle):
        self.L, self.h, self._keep = lib, C.c_void_p(handle), []

    def close(self):
        self.h = C.c_void_p()


class SlabbedSim:
    """vx_slabbed_*: the WHOLE model in the caller's numbering, run as z-slabs on the listed devices of this process
    (include/voxelyze_b200.h).  Same method names and meaning as Sim where they exist."""

    def __init__(self, lib: VxLib, voxel_size: float, devices: Sequence[int]):
        self.L = lib
        self.h = C.c_void_p()
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        rc = lib.lib.vx_slabbed_create(float(voxel_size), len(dev), _ptr(dev), C.byref(self.h))
        if rc != VX_OK:
            raise VxError(rc, f"vx_slabbed_create failed on {lib.path}")
        self.voxel_size = voxel_size

    def close(self):
        if self.h:
            self.L.lib.vx_slabbed_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc: int, ok=(VX_OK,)):
        if rc not in ok:
            raise VxError(rc, self.L.lib.vx_slabbed_last_error(self.h).decode())
        return rc

    @property
    def n_slabs(self) -> int:
        return self.L.lib.vx_slabbed_slab_count(self.h)

    @property
    def halo_mode(self) -> int:
        return self.L.lib.vx_slabbed_halo_mode(self.h)

    def slab(self, k: int) -> Sim:
        return _SlabView(self.L, self.L.lib.vx_slabbed_slab(self.h, k))

    def set_materials(self, mats: Sequence[Material]):
        arr, keep = _material_descs(mats)
        self._chk(self.L.lib.vx_slabbed_set_materials(self.h, len(mats), arr))

    def set_gravity(self, g: float):
        self._chk(self.L.lib.vx_slabbed_set_gravity(self.h, g))

    def enable_floor(self, on: bool = True):
        self._chk(self.L.lib.vx_slabbed_enable_floor(self.h, int(on)))

    def set_voxels(self, ijk, mat):
        ijk = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        mat = np.ascontiguousarray(mat, dtype=np.uint16)
        self._chk(self.L.lib.vx_slabbed_set_voxels(self.h, len(ijk), _ptr(ijk), _ptr(mat)))

    @property
    def n_voxels(self) -> int:
        return self.L.lib.vx_slabbed_voxel_count(self.h)

    @property
    def n_links(self) -> int:
        return self.L.lib.vx_slabbed_link_count(self.h)

    def links(self):
        n = self.n_links
        vn, vp, ax = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
        self._chk(self.L.lib.vx_slabbed_get_links(self.h, _ptr(vn), _ptr(vp), _ptr(ax)))
        return vn, vp, ax

    def set_externals(self, voxel, dof, force=None, moment=None, translation=None, rotation=None):
        voxel = np.ascontiguousarray(voxel, dtype=np.int32)
        n = len(voxel)
        dof = np.ascontiguousarray(dof, dtype=np.uint8)
        f = None if force is None else np.ascontiguousarray(force, dtype=np.float32).reshape(n, 3)
        m = None if moment is None else np.ascontiguousarray(moment, dtype=np.float32).reshape(n, 3)
        t = None if translation is None else np.ascontiguousarray(translation, dtype=np.float64).reshape(n, 3)
        r = None if rotation is None else np.ascontiguousarray(rotation, dtype=np.float64).reshape(n, 3)
        self._chk(self.L.lib.vx_slabbed_set_externals(self.h, n, _ptr(voxel), _ptr(dof), _ptr(f), _ptr(m), _ptr(t), _ptr(r)))

    def set_temperature_all(self, t: float):
        self._chk(self.L.lib.vx_slabbed_set_temperature_all(self.h, t))

    def set_temperature(self, t):
        t = np.ascontiguousarray(t, dtype=np.float32)
        self._chk(self.L.lib.vx_slabbed_set_temperature(self.h, len(t), _ptr(t)))

    def step(self, dt: float, n: int = 1) -> Optional[int]:
        div = C.c_int(-1)
        rc = self._chk(self.L.lib.vx_slabbed_step(self.h, dt, n, C.byref(div)), ok=(VX_OK, VX_DIVERGED))
        return div.value if rc == VX_DIVERGED else None

    def recommended_dt(self) -> float:
        dt = C.c_float()
        self._chk(self.L.lib.vx_slabbed_recommended_dt(self.h, C.byref(dt)))
        return dt.value

    def reset(self):
        self._chk(self.L.lib.vx_slabbed_reset(self.h))

    def time(self) -> float:
        return self.L.lib.vx_slabbed_time(self.h)

    def set_clock(self, time: float, previous_dt: float):
        self._chk(self.L.lib.vx_slabbed_set_clock(self.h, time, previous_dt))

    def download(self, name: str, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        fid, dt, comps, is_link = FIELDS[name]
        total = self.n_links if is_link else self.n_voxels
        if count is None:
            count = total - first
        out = np.zeros((count, comps), dtype=dt)
        if count:
            self._chk(self.L.lib.vx_slabbed_download(self.h, fid, first, count, _ptr(out)))
        return out if comps > 1 else out.reshape(-1)

    def upload(self, name: str, data, first: int = 0):
        fid, dt, comps, _ = FIELDS[name]
        a = np.ascontiguousarray(data, dtype=dt).reshape(-1, comps)
        self._chk(self.L.lib.vx_slabbed_upload(self.h, fid, first, len(a), _ptr(a)))

    def download_voxel_state(self, first: int = 0, count: int = 1) -> np.ndarray:
        out = np.zeros(count, Sim.VOXEL_STATE_DTYPE)
        self._chk(self.L.lib.vx_slabbed_download_voxel_state(self.h, first, count, out.ctypes.data))
        return out

    def download_link_state(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.n_links - first if count is None else count
        out = np.zeros(count, Sim.LINK_STATE_DTYPE)
        self._chk(self.L.lib.vx_slabbed_download_link_state(self.h, first, count, out.ctypes.data))
        return out

    def upload_link_state(self, rec: np.ndarray, first: int = 0):
        rec = np.ascontiguousarray(rec, dtype=Sim.LINK_STATE_DTYPE)
        self._chk(self.L.lib.vx_slabbed_upload_link_state(self.h, first, len(rec), rec.ctypes.data))

    def state_info(self, info: int, typ: int) -> float:
        v = C.c_float()
        self._chk(self.L.lib.vx_slabbed_state_info(self.h, info, typ, C.byref(v)))
        return v.value

    def save_state(self, path: str):
        self._chk(self.L.lib.vx_slabbed_save_state(self.h, os.fsencode(path)))

    def load_state(self, path: str):
        self._chk(self.L.lib.vx_slabbed_load_state(self.h, os.fsencode(path)))

    def launch_count(self) -> int:
        return self.L.lib.vx_slabbed_launch_count(self.h)


_cache: dict = {}


def _load(path: str) -> VxLib:
    if path not in _cache:
        _cache[path] = VxLib(path)
    return _cache[path]


def load_product() -> VxLib:
    """The CUDA library.  Raises if it has not been built; never substitutes a CPU path."""
    return _load(os.environ.get("VX_PRODUCT_SO", PRODUCT_SO))   # override: kernel-variant experiments only


def load_oracle() -> VxLib:
    """CPU restatement -- tests / smoke / bench cpu_baseline only."""
    return _load(ORACLE_SO)


def load_reference(omp: bool = False) -> VxLib:
    """The unmodified reference behind oracle/ref_shim.cpp -- tests / bench baseline only."""
    return _load(REF_OMP_SO if omp else REF_SO)
