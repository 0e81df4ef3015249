"""Shared helpers for parity tests: run the same scenario on two C-ABI libraries and compare."""
from __future__ import annotations

import numpy as np

from voxelyze_b200 import capi, scenarios

VOXEL_FIELDS = ["pos", "orient", "linmom", "angmom", "temp", "voxflags"]
LINK_FIELDS = ["force_neg", "force_pos", "moment_neg", "moment_pos", "pos2", "angle1v", "angle2v",
               "strain", "maxstrain", "strainoffset", "stress", "linkflags"]


def snapshot(sim: capi.Sim, fields=None) -> dict:
    fields = fields or (VOXEL_FIELDS + LINK_FIELDS)
    return {f: sim.download(f) for f in fields}


def run(lib: capi.VxLib, sc: scenarios.Scenario, steps: int, dt=None, program=None, chunk=None, device=0, path=0):
    """Builds `sc`, steps it, returns (sim, dt, diverged_at).  `program(sim, k, t)` is called
    before step k (for per-step temperature / force changes); with a program steps run one by one."""
    sim = scenarios.build(lib, sc, device, path)
    if dt is None:
        dt = sc.dt if sc.dt is not None else sim.recommended_dt()
    div = None
    if program is None:
        div = sim.step(dt, steps)
    else:
        t = 0.0
        for k in range(steps):
            program(sim, k, t)
            div = sim.step(dt, 1)
            if div is not None:
                div = k
                break
            t = float(np.float32(t) + np.float32(dt))
    return sim, dt, div


SMALL_MAX = 700     # VX_SMALL_MAX of csrc/vx_capi.cu: below it "auto" steps the general layout with k_small_steps


def layout(path: int, n_voxels: int, collisions: bool = False) -> int:
    """vx_active_path() a model is expected to report: 1 general layout, 2 fused lattice layout."""
    if path in (1, 3):
        return 1
    if path == 0 and n_voxels <= SMALL_MAX and not collisions:
        return 1
    return 2


def bit_equal(a: np.ndarray, b: np.ndarray) -> bool:
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    # -0.0 == +0.0 is accepted (a missing external adds no +0.0, see ref_shim.cpp); NaNs must match
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b)))) if a.dtype.kind == "f" else bool(np.array_equal(a, b))


def displacement_scale(sim: capi.Sim, sc: scenarios.Scenario) -> float:
    pos = sim.download("pos")
    nominal = sc.ijk.astype(np.float64) * sc.voxel_size
    return float(np.max(np.abs(pos - nominal)))


def rel_errors(got: dict, ref: dict, sc: scenarios.Scenario) -> dict:
    """Parity metric of SURVEY.md section 8d: max |p - p_ref|_inf / max |p_ref - p_nominal|_inf
    for positions; absolute component error for quaternions; momenta/forces scaled by their max."""
    nominal = sc.ijk.astype(np.float64) * sc.voxel_size
    out = {}
    dscale = max(float(np.max(np.abs(ref["pos"] - nominal))), 1e-300)
    out["pos"] = float(np.max(np.abs(got["pos"] - ref["pos"]))) / dscale
    out["orient"] = float(np.max(np.abs(got["orient"] - ref["orient"])))
    for f in ("linmom", "angmom", "force_neg", "force_pos", "moment_neg", "moment_pos", "pos2", "angle1v", "angle2v",
              "strain", "stress", "maxstrain", "strainoffset"):
        if f in got and f in ref and ref[f].size:
            sc_ = max(float(np.max(np.abs(ref[f]))), 1e-300)
            out[f] = float(np.max(np.abs(got[f].astype(np.float64) - ref[f].astype(np.float64)))) / sc_
    return out
