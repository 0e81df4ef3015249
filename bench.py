#!/usr/bin/env python
"""bench.py -- link+voxel updates/s of the explicit dynamics step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # ours (CUDA, sm_100a)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path

A "step" is one CVoxelyze::doTimeStep over the whole lattice = N_vox voxel integrations +
N_link link force evaluations ("updates").  Workloads:
  N = 1 : C5a, 256^3 solid cantilever (16 777 216 voxels, 50 135 040 links), resident in HBM
  N > 1 : C5b, 512^3 cantilever split into z-slabs, one per GPU, one-voxel pose halo
          exchanged every step over NVLink (NCCL send/recv); strong scaling of one lattice.
Inputs are far larger than L2 (>= 10 GB of state per step vs 126 MB), so no L2 flush is needed.
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


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def ncu_traffic(kernel: str, voxels: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json); None when no capture matches this kernel and size."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            for row in json.load(f):
                if kernel.startswith(row["kernel"]) and row["voxels"] == voxels:
                    return row["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_arm(args, sample_n: int, emit: bool):
    """Times the reference's own CPU implementation (oracle/_ref, OpenMP build, all host threads)
    on a bounded sample of the workload: an n^3 cantilever of the same pattern."""
    from voxelyze_b200 import capi, scenarios
    kind, cores = "reference", os.cpu_count() or 1
    if os.path.exists(capi.REF_OMP_SO):
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            os.environ["OMP_NUM_THREADS"] = str(cores)      # torchrun pins it to 1; rank 0 is the only rank that computes here
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        os.environ.setdefault("OMP_PROC_BIND", "close")
        lib = capi.load_reference(omp=True)
    elif os.path.exists(capi.REF_SO):
        lib, cores = capi.load_reference(), 1
    else:
        lib, kind, cores = capi.load_oracle(), "port", 1
    sc = scenarios.cantilever(sample_n, sample_n, sample_n, tip_load=1.0)
    sim = scenarios.build(lib, sc)
    dt = sim.recommended_dt()
    units = sim.n_voxels + sim.n_links
    sim.step(dt, max(args.warmup, 1))
    per = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        sim.step(dt, 1)
        per.append(time.perf_counter() - t0)
    total = sum(per)
    value = units * len(per) / total
    sample = f"{sample_n}^3 cantilever (same pattern as the workload), {len(per)} steps, {lib.backend}"
    base = {"value": value, "unit": "updates/s", "cores": cores, "kind": kind, "sample": sample}
    if emit:
        line = {"impl": "reference", "metric": "link+voxel updates/sec", "value": value, "unit": "updates/s",
                "n_gpus": args.gpus, "steps": len(per), "warmup": args.warmup, "ms_per_step": 1e3 * total / len(per),
                "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args.gpus), "sample": sample},
                "cpu_baseline": base,
                "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
    return base


def workload_name(gpus: int) -> str:
    return ("C5a: 256^3 solid cantilever, x=0 face fixed, -1/65536 N on each x=255 face voxel, E=1e6 rho=1e3"
            if gpus == 1 else
            f"C5b: 512^3 solid cantilever in {gpus} z-slabs with one-voxel pose halo exchange per step")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=0, help="override lattice edge (testing only; reported in config)")
    ap.add_argument("--cpu-sample", type=int, default=128, help="edge of the CPU baseline sample lattice")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peer", action="store_true", help="N>1: NCCL send/recv for the halo instead of peer-memory stores")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: exchange the halo after the step instead of overlapping it with the interior")
    ap.add_argument("--path", type=int, default=0, help="kernel variant (vx_set_path): 0 auto, 1 general, 2..5 fused lattice variants (ablation)")
    args = ap.parse_args()
    # >= 17 (one direct step + one 16-step graph) so that the CUDA graphs are captured and instantiated before the timed region
    args.warmup = max(args.warmup, 20) if args.impl == "ours" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank == 0:
            cpu_reference_arm(args, args.cpu_sample, emit=True)
        return

    import torch
    import torch.distributed as dist
    from voxelyze_b200 import capi, scenarios, slab

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.load_product()          # raises when the CUDA library is missing: no CPU fallback

    edge = args.size or (256 if world == 1 else 512)
    stream = torch.cuda.current_stream()
    if world == 1:
        sc = scenarios.cantilever(edge, edge, edge, tip_load=1.0)
        sim = scenarios.build(lib, sc, device=local, path=args.path)
        runner = slab.SingleRunner(sim)
    else:
        runner = slab.SlabRunner(lib, edge, edge, edge, rank, world, device=local, path=args.path, overlap=not args.no_overlap, peer=not args.no_peer)
        sim = runner.sim
    sim.set_stream(stream.cuda_stream)
    dt = runner.recommended_dt()
    n_vox, n_link = runner.global_counts()
    units = n_vox + n_link

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    runner.step(dt, args.warmup)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
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

    # ---- per-kernel roofline, live CUDA events around the dominant kernel (link-force kernels)
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
    launch_bytes = (B_LINK * own_link + (B_VOXEL * own_vox if fused else 0)) / max(kl[0] // prof_steps, 1)
    launch_ms = link_ms / max(kl[0] // prof_steps, 1)
    achieved = launch_bytes / (launch_ms * 1e-3) / 1e9 if link_ms > 0 else 0.0
    step_gbs = (B_VOXEL * n_vox + B_LINK * n_link) * args.steps / (ms * 1e-3) / 1e9 / world
    roofline = {"bound": "hbm", "kernel": runner.dominant_kernel(), "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(runner.dominant_kernel(), own_vox),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": launch_bytes, "launch_ms": launch_ms,
                "launches_per_step": kl[0] // prof_steps, **({"kernel_ms_per_rank": per_rank_ms} if world > 1 else {}),
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
    e2e = {"value": units * e2e_steps / e2e_s, "unit": "updates/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 32 + 24,
           "note": "per-step blocking vx_step(dt,1) + vx_download of one voxel position; lattice state stays in HBM "
                   "like the reference keeps it in RAM (construction excluded on both arms)"}

    if rank == 0:
        line = {"metric": "link+voxel updates/sec", "value": value, "unit": "updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(world) if not args.size else f"{edge}^3 cantilever (size override)",
                           "voxels": n_vox, "links": n_link, "dt": dt, "path": runner.path_name(), **({"halo": runner.halo_name()} if world > 1 else {}),
                           "l2": "inputs larger than L2 (no flush needed)"},
                "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline}
        if not args.no_cpu_baseline and world == 1:          # the CPU baseline is reported at N = 1 only (other ranks would wait on it)
            cb_args = argparse.Namespace(**vars(args)); cb_args.steps, cb_args.warmup = 10, 2
            line["cpu_baseline"] = cpu_reference_arm(cb_args, args.cpu_sample, emit=False)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
