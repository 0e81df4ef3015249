#!/usr/bin/env python
"""bench.py -- link+voxel updates/s of the explicit dynamics step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c5|c4]      # ours (CUDA, sm_100a)
    python bench.py --impl reference [--steps K] [--warmup W] [--config ...]  # the reference's own CPU code

A "step" is one CVoxelyze::doTimeStep over the whole workload = N_vox voxel integrations + N_link link force
evaluations ("updates").  Workloads:
  c5 (default)  N = 1: C5a, 256^3 solid cantilever (16 777 216 voxels, 50 135 040 links)
                N > 1: C5b, 512^3 cantilever split into z-slabs, one per GPU, one-voxel pose halo pushed into the
                       neighbours' ghost layers over NVLink by the step kernel itself (strong scaling of ONE lattice)
  c4            C4, 4096 soft robots of 10^3 voxels, three materials, CTE actuation set every step, floor + gravity;
                4096/N robots per GPU, no communication (strong scaling of one population)
Inputs are far larger than L2 (>= 2 GB of state per step against 126 MB), so no L2 flush is needed.

Correctness travels with the number (outside every timed region):
  N = 1  all voxel poses after the first PARITY_STEPS steps are compared with the unmodified reference's OpenMP build
         run on the SAME workload on the host (the same run is the `cpu_baseline`): `parity`, exit code 3 above 1e-9
  N > 1  the owned state of every rank after the same steps is compared, z-plane by z-plane and bit for bit, with a
         one-GPU run of the whole 512^3 lattice on rank 0: `slab_bitwise`, exit code 3 if false
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_VOXEL, B_LINK = 216, 268          # algorithmic bytes per voxel / link update
FALLBACK_HBM_GBS = 6650.0           # /opt/skills/guides/B200_PROFILING.md fallback
PARITY_STEPS = 20                   # steps of the correctness legs (verdict r1: "the same 20 steps")
PARITY_TOL = 1e-9                   # north_star tolerance on smooth cases


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def ncu_traffic(kernel: str, voxels: int):
    """(dram bytes per launch, source) of the dominant kernel from the committed `ncu --set full` captures
    (profiles/ncu_traffic.json) -- a looked-up constant of an earlier capture of the same kernel and size, NOT measured
    in this run (ncu cannot run inside a timed bench); (None, None) when no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            for row in json.load(f):
                if kernel.startswith(row["kernel"]) and row["voxels"] == voxels:
                    return row["dram_bytes_per_launch"], "profiles/" + row.get("capture", "ncu_traffic.json")
    except Exception:
        pass
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled from before the warm-up to the end of the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.t_timed = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_timed(self):
        self.t_timed = time.perf_counter()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons, busy = [], [], set(), []
        for t, r in self.rows:
            try:
                c, m, w = float(r[1]), float(r[2]), float(r[3])
            except Exception:
                continue
            sm.append(c); mx.append(m)
            if self.t_timed is not None and t >= self.t_timed:
                busy.append(c)
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        under = busy or sm
        return {"sm_mhz": statistics.median(under) if under else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": len(busy),
                "window": "from before the warm-up to the end of the timed region"}


# ------------------------------------------------------------------------------------------------
# workloads
def workload_name(config: str, gpus: int) -> str:
    if config == "c4":
        return ("C4: 4096 soft robots of 10^3 voxels, 3 materials (CTE +0.01/-0.01/0), floor + gravity, ambient "
                f"temperature 20 sin(2 pi 40 t) set every step; {4096 // gpus} robots per GPU, no communication")
    return ("C5a: 256^3 solid cantilever, x=0 face fixed, -1/65536 N on each x=255 face voxel, E=1e6 rho=1e3"
            if gpus == 1 else
            f"C5b: 512^3 solid cantilever in {gpus} z-slabs with one-voxel pose halo exchange per step")


def workload_counts(config: str, gpus: int, edge: int = 0):
    if config == "c4":
        return 4096 * 1000, 4096 * 3 * 10 * 10 * 9
    n = edge or (256 if gpus == 1 else 512)
    return n ** 3, 3 * n * n * (n - 1)


def config_block(config: str, gpus: int, dt: float, edge: int = 0) -> dict:
    """Identical in both arms (the driver compares it)."""
    v, l = workload_counts(config, gpus, edge)
    return {"workload": workload_name(config, gpus) if not edge else f"{edge}^3 cantilever (size override)",
            "voxels": v, "links": l, "dt": dt, "l2": "inputs larger than L2 (no flush needed)"}


def physical_cores() -> int:
    try:
        import psutil
        return psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        return os.cpu_count() or 1


def load_cpu_lib():
    """The reference's own code (oracle/_ref, OpenMP build) on all physical cores; falls back to the serial build and
    then to the oracle port.  Returns (lib, kind, threads)."""
    from voxelyze_b200 import capi
    if os.path.exists(capi.REF_OMP_SO):
        cores = physical_cores()
        os.environ["OMP_NUM_THREADS"] = str(cores)          # torchrun pins it to 1; rank 0 is the only rank that computes here
        os.environ["OMP_PLACES"] = "cores"
        os.environ["OMP_PROC_BIND"] = "close"
        return capi.load_reference(omp=True), "reference", cores
    if os.path.exists(capi.REF_SO):
        return capi.load_reference(), "reference", 1
    return capi.load_oracle(), "port", 1


def cpu_scenario(config: str, gpus: int, edge: int = 0):
    """What the CPU arm runs: the workload itself when it fits a host (C5a), else a bounded sample of it."""
    from voxelyze_b200 import scenarios
    if config == "c4":
        return scenarios.robot_ensemble(64, 10), "64 of the 4096 robots (seeds 0..63), temperature set every step", False
    n = edge or 256
    whole = gpus == 1
    what = (f"the whole workload ({n}^3)" if whole else f"{n}^3 cantilever of the same pattern (1/8 of the 512^3 workload's voxels; "
            "the whole lattice needs ~160 GB of host memory in the reference's object graph)")
    return scenarios.cantilever(n, n, n, tip_load=1.0), what, whole


def cpu_run(config: str, gpus: int, steps: int, warmup: int, edge: int = 0, keep_state: bool = False):
    """Builds the CPU scenario on the reference and times `steps` steps one by one after `warmup` untimed ones.
    Returns (baseline dict, dt, state or None)."""
    from voxelyze_b200 import scenarios
    lib, kind, cores = load_cpu_lib()
    sc, what, whole = cpu_scenario(config, gpus, edge)
    t0 = time.perf_counter()
    sim = scenarios.build(lib, sc)
    build_s = time.perf_counter() - t0
    dt = sim.recommended_dt()
    units = sim.n_voxels + sim.n_links
    per = []

    def one():
        if config == "c4":
            sim.set_temperature_all(scenarios.robot_temperature(sim.time()))
        sim.step(dt, 1)
    for _ in range(warmup):
        one()
    for _ in range(steps):
        t = time.perf_counter()
        one()
        per.append(time.perf_counter() - t)
    total = sum(per)
    base = {"value": units * len(per) / total, "unit": "updates/s", "cores": cores, "kind": kind,
            "sample": f"{what}: {sim.n_voxels} voxels + {sim.n_links} links, {len(per)} timed steps after {warmup}, {lib.backend}, "
                      f"OMP_NUM_THREADS={cores} (physical cores) OMP_PLACES=cores OMP_PROC_BIND=close",
            "same_workload": bool(whole), "ms_per_step": 1e3 * total / len(per), "build_s": round(build_s, 1)}
    state = None
    if keep_state:
        state = {"pos": sim.download("pos"), "orient": sim.download("orient"), "steps": warmup + steps}
    sim.close()
    return base, dt, state


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path, all physical cores, same config block."""
    base, dt, _ = cpu_run(args.config, args.gpus, args.steps, max(args.warmup, 1), args.size)
    line = {"impl": "reference", "metric": "link+voxel updates/sec", "value": base["value"], "unit": "updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": base["ms_per_step"],
            "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_block(args.config, args.gpus, dt, args.size),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# correctness legs
def plane_checksums(download, first_plane: int, n_planes: int, plane: int, chunk_planes: int = 16) -> dict:
    """Per z-plane 64-bit checksums (wrapping sum and xor of the bit patterns, -0.0 folded into +0.0) of the four
    voxel state arrays; `download(field, first, count)` addresses voxels plane-major."""
    out = {}
    for f in ("pos", "orient", "linmom", "angmom"):
        sums, xors = [], []
        for p0 in range(0, n_planes, chunk_planes):
            k = min(chunk_planes, n_planes - p0)
            a = download(f, (first_plane + p0) * plane, k * plane)
            a = (a + 0.0).reshape(k, -1).view(np.uint64)
            sums.append(np.add.reduce(a, axis=1, dtype=np.uint64))
            xors.append(np.bitwise_xor.reduce(a, axis=1))
        out[f] = (np.concatenate(sums), np.concatenate(xors))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c5", choices=["c5", "c4"])
    ap.add_argument("--size", type=int, default=0, help="c5: override lattice edge (testing only; reported in config)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline + parity leg (N=1) / the whole-lattice check (N>1)")
    ap.add_argument("--no-facade", action="store_true", help="skip the e2e leg through the C++ facade")
    ap.add_argument("--no-peer", action="store_true", help="N>1: NCCL send/recv for the halo instead of peer-memory stores")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: exchange the halo after the step instead of overlapping it with the interior")
    ap.add_argument("--path", type=int, default=0, help="kernel variant (vx_set_path): 0 auto, 1 general, 5/7 fused lattice with cp.async/TMA staging")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return 0

    import torch
    import torch.distributed as dist
    from voxelyze_b200 import capi, scenarios, slab

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.load_product()          # raises when the CUDA library is missing: no CPU fallback
    clocks = ClockSampler(local)
    clocks.start()                     # before the warm-up, so that short timed regions still collect samples

    stream = torch.cuda.current_stream()
    c4 = args.config == "c4"
    if c4:
        robots = 4096 // world
        sc = scenarios.robot_ensemble(robots, 10, first_seed=rank * robots)
        sim = scenarios.build(lib, sc, device=local, path=args.path)
        runner = slab.EnsembleRunner(sim, world)
    elif world == 1:
        edge = args.size or 256
        sc = scenarios.cantilever(edge, edge, edge, tip_load=1.0)
        sim = scenarios.build(lib, sc, device=local, path=args.path)
        runner = slab.SingleRunner(sim)
    else:
        edge = args.size or 512
        runner = slab.SlabRunner(lib, edge, edge, edge, rank, world, device=local, path=args.path, overlap=not args.no_overlap, peer=not args.no_peer)
        sim = runner.sim
    sim.set_stream(stream.cuda_stream)
    dt = runner.recommended_dt()
    n_vox, n_link = runner.global_counts()
    units = n_vox + n_link
    verify = not args.no_cpu_baseline

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness leg, GPU half: the first PARITY_STEPS steps from the fresh state (untimed, before the warm-up)
    gpu_state, sums = None, None
    if verify and not c4:
        runner.step(dt, PARITY_STEPS)
        torch.cuda.synchronize()
        if world == 1:
            gpu_state = {"pos": sim.download("pos"), "orient": sim.download("orient")}
        else:
            first, _ = runner.layer_index_range(runner.z0)
            sums = plane_checksums(lambda f, a, n: sim.download(f, first + a, n), 0, runner.z1 - runner.z0, runner.plane)
    sim.prepare()                      # captured step graphs built now, not inside the warm-up

    runner.step(dt, args.warmup)
    barrier()
    l0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.mark_timed()
    e0.record(stream)
    runner.step(dt, args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sim.launch_count() - l0
    clk = clocks.stop()
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        t = torch.tensor([launches], device="cuda", dtype=torch.int64); dist.all_reduce(t); launches = int(t.item())
    value = units * args.steps / (ms * 1e-3)

    # ---- per-kernel roofline, live CUDA events around the dominant kernel
    prof_steps = min(args.steps, 20)
    kms, kl = runner.step_profile(dt, prof_steps)
    own_vox, own_link = runner.local_counts()
    peak, peak_src = measured_hbm_peak()
    link_ms = kms["link"] / prof_steps
    per_rank_ms = [link_ms]
    if world > 1:                       # the slabs are coupled through the halo: the slowest GPU sets the pace
        t = torch.zeros(world, device="cuda"); t[rank] = link_ms
        dist.all_reduce(t); per_rank_ms = [round(float(x), 4) for x in t.tolist()]
    fused = kl[1] == 0                  # lattice path: the one kernel does the link AND the voxel updates
    per_step_launches = max(kl[0] // prof_steps, 1)
    launch_bytes = (B_LINK * own_link + (B_VOXEL * own_vox if fused else 0)) / per_step_launches
    launch_ms = link_ms / per_step_launches
    achieved = launch_bytes / (launch_ms * 1e-3) / 1e9 if link_ms > 0 else 0.0
    step_gbs = (B_VOXEL * n_vox + B_LINK * n_link) * args.steps / (ms * 1e-3) / 1e9 / world
    traffic, traffic_src = ncu_traffic(runner.dominant_kernel(), own_vox)
    roofline = {"bound": "hbm", "kernel": runner.dominant_kernel(), "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src and f"{traffic_src}: ncu --set full capture of this kernel at this size, not measured in this run",
                "dram_frac": (traffic / (launch_ms * 1e-3) / 1e9 / peak) if traffic and launch_ms > 0 else None,
                "note": "achieved/frac use the ALGORITHMIC bytes of SURVEY 8d (216 B/voxel + 268 B/link); the kernel's real DRAM traffic "
                        "is `traffic` (dram_frac = traffic / launch time / peak): the fused kernel moves ~64 % of the algorithmic bytes "
                        "and is FP64-issue/latency bound, not bandwidth bound",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": launch_bytes, "launch_ms": launch_ms,
                "launches_per_step": per_step_launches, **({"kernel_ms_per_rank": per_rank_ms} if world > 1 else {}),
                "kernel_ms_per_step": {k: v / prof_steps for k, v in kms.items()},
                "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak,
                               "bytes_per_step": B_VOXEL * n_vox + B_LINK * n_link}}

    # ---- e2e: the call a user of the reference makes every step (test/tVoxelyze.h:94-107):
    # doTimeStep(dt) as one blocking C-ABI call (dt goes host->device, the divergence/status block
    # comes back) followed by a position() read of one voxel into a host buffer.
    e2e_steps = min(args.steps, 50)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        runner.step(dt, 1)
        runner.read_probe()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
    e2e = {"value": units * e2e_steps / e2e_s, "unit": "updates/s", "h2d_bytes_per_step": runner.h2d_bytes_per_step(), "d2h_bytes_per_step": 40 + 24,
           "note": "per-step blocking vx_step(dt,1) + vx_download of one voxel position through the C-ABI with host buffers; lattice state "
                   "stays in HBM like the reference keeps it in RAM (construction excluded on both arms)"}

    # ---- the same per-step loop through the C++ class API of the facade (CVoxelyze::doTimeStep + CVX_Voxel::position), the
    # binding a drop-in caller of the reference actually uses; separate process, same GPU, after the device-timed region
    if world == 1 and not c4 and not args.no_facade:
        exe = os.path.join(ROOT, "tests", "cpp", "_build", "facade_e2e")
        try:
            out = subprocess.run([exe, str(args.size or 256), "5", str(e2e_steps)], capture_output=True, text=True, timeout=600)
            fe = json.loads(out.stdout.strip().splitlines()[-1])
            # the headline e2e is the one through the reference-facing API itself: the C++ class API a drop-in caller links against
            e2e = {"value": fe["updates_per_s"], "unit": "updates/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 40 + 112,
                   "ms_per_step": fe["ms_per_step"], "steps": fe["steps"], "build_s": fe["build_s"],
                   "note": "CVoxelyze::doTimeStep(dt) + CVX_Voxel::position() of one voxel per step through libvoxelyze_facade.so (tools/facade_e2e.cpp, "
                           "a separate process on the same GPU): dt down, the 40-byte status block and one 112-byte voxel record up, every step, "
                           "blocking; lattice state stays in HBM like the reference keeps it in RAM (construction excluded on both arms)",
                   "capi_ctypes": {"value": e2e["value"], "unit": "updates/s", "note": "the same loop through the C-ABI from Python (ctypes): vx_step(dt,1) + vx_download"}}
        except Exception as exc:                        # missing binary (not built) or a failed run: keep the ctypes number and say so
            e2e["facade_error"] = f"{type(exc).__name__}: {exc}"[:200]

    line = {"metric": "link+voxel updates/sec", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(args.config, world, dt, args.size),
            "detail": {"path": runner.path_name(), **({"halo": runner.halo_name()} if world > 1 and not c4 else {}),
                       "untimed_steps_before_warmup": PARITY_STEPS if (verify and not c4) else 0,
                       "graphs": "16-step CUDA graphs instantiated by vx_prepare before the warm-up"},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline}
    rc = 0

    # ---- correctness leg, reference half (outside every timed region)
    if verify and world == 1 and not c4:
        base, dt_cpu, ref = cpu_run(args.config, 1, PARITY_STEPS - 2, 2, args.size, keep_state=True)    # 2 untimed + 18 timed = the same 20 steps
        line["cpu_baseline"] = base
        nominal = sc.ijk.astype(np.float64) * sc.voxel_size
        dscale = max(float(np.max(np.abs(ref["pos"] - nominal))), 1e-300)
        perr = float(np.max(np.abs(gpu_state["pos"] - ref["pos"]))) / dscale
        oerr = float(np.max(np.abs(gpu_state["orient"] - ref["orient"])))
        ok = perr <= PARITY_TOL and oerr <= PARITY_TOL and np.float32(dt_cpu) == np.float32(dt)
        line["parity"] = {"pos": perr, "orient": oerr, "steps": PARITY_STEPS, "voxels_compared": int(len(nominal)), "tol": PARITY_TOL,
                          "dt_equal": bool(np.float32(dt_cpu) == np.float32(dt)), "ok": bool(ok),
                          "against": base["sample"], "metric": "max|p-p_ref|_inf / max|p_ref-p_nominal|_inf; max quaternion component difference"}
        rc = 0 if ok else 3
    elif verify and c4 and rank == 0:
        base, _, _ = cpu_run("c4", world, 18, 2)
        line["cpu_baseline"] = base
    elif verify and world > 1:
        gathered = [None] * world
        dist.gather_object((runner.z0, runner.z1, sums), gathered if rank == 0 else None, dst=0)
        runner.close()                                      # unmap the neighbours' ghost layers, free the slab
        barrier()
        if rank == 0:
            t0 = time.perf_counter()
            whole_sc = scenarios.cantilever(edge, edge, edge, tip_load=1.0)
            whole = scenarios.build(lib, whole_sc, device=local, path=args.path)
            whole.step(dt, PARITY_STEPS)
            ref = plane_checksums(lambda f, a, n: whole.download(f, a, n), 0, edge, edge * edge)
            whole.close()
            bad = []
            for z0, z1, s in gathered:
                for f, (ssum, sxor) in s.items():
                    if not (np.array_equal(ssum, ref[f][0][z0:z1]) and np.array_equal(sxor, ref[f][1][z0:z1])):
                        planes = np.nonzero((ssum != ref[f][0][z0:z1]) | (sxor != ref[f][1][z0:z1]))[0] + z0
                        bad.append((f, planes[:4].tolist()))
            line["slab_bitwise"] = not bad
            line["slab_check"] = {"steps": PARITY_STEPS, "planes": edge, "fields": ["pos", "orient", "linmom", "angmom"],
                                  "against": f"one-GPU run of the whole {edge}^3 lattice on rank 0 (same steps, same dt)",
                                  "mismatches": bad[:8], "seconds": round(time.perf_counter() - t0, 1)}
            rc = 0 if not bad else 3
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
