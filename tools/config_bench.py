#!/usr/bin/env python
"""Throughput of the BASELINE.json configurations other than the headline one.

    python tools/config_bench.py [--config c1,c2,c3,c4] [--impl ours|reference] [--steps K] [--scale S]
    torchrun --nproc-per-node N tools/config_bench.py --config c4        # ensemble sharded over N GPUs

  C1  20x4x4 cantilever                         (launch-bound: what CUDA graphs are for)
  C2  64^3 block dropped on the floor           (floor contact, gravity; fused lattice path)
  C3  128x128 plate stack, 16 plates            (two bilinear materials, self collisions; general path)
  C4  4096 robots of 10^3, CTE actuation        (ensemble; temperature set from the host EVERY step)
`--impl reference` runs the unmodified reference (oracle/_ref, OpenMP) on a bounded sample of the same
configuration (`--scale` divides the size) so that it finishes in seconds.  One JSON line per config.
Metric: (voxels + links) * steps / seconds, wall clock around the blocking calls a user makes.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelyze_b200 import capi, scenarios  # noqa: E402


def make(config: str, scale: int, rank: int, world: int):
    if config == "c1":
        return scenarios.cantilever(), "20x4x4 cantilever"
    if config == "c2":
        n = max(64 // scale, 4)
        return scenarios.drop_block(n), f"{n}^3 block dropped one voxel onto the floor"
    if config == "c3":
        n, plates = max(128 // scale, 16), max(16 // scale, 2)
        return scenarios.plate_stack(n, n, plates), f"{plates} plates of {n}x{n}x6, 4^3 checkerboard of two bilinear materials, collisions on"
    if config == "c4":
        robots = max(4096 // scale, 8) // world
        return scenarios.robot_ensemble(robots, 10, first_seed=rank * robots), f"{robots * world} robots of 10^3 ({robots} per GPU), CTE actuation set every step"
    raise SystemExit(f"unknown config {config}")


def run(lib, config: str, args, rank: int, world: int, sync):
    sc, what = make(config, args.scale, rank, world)
    sim = scenarios.build(lib, sc, device=int(os.environ.get("LOCAL_RANK", "0")) if lib.backend.startswith("cuda") else 0)
    dt = sim.recommended_dt()
    units = sim.n_voxels + sim.n_links
    per_step_host = config == "c4"

    def advance(n):
        if per_step_host:                                  # src: per-step setAmbientTemperature(20 sin(2 pi 40 t))
            for _ in range(n):
                sim.set_temperature_all(scenarios.robot_temperature(sim.time()))
                sim.step(dt, 1)
        else:
            sim.step(dt, n)

    # C3 is timed with LIVE contacts: under gravity the cantilever plates sag onto the plate that rests on the floor (first contact
    # after ~5 700 steps at full size), so its warm-up runs until contacts exist; pairs and rebuilds before/after are reported
    warm = max(args.warmup, args.c3_warmup // max(args.scale, 1)) if config == "c3" else args.warmup
    advance(warm)
    sync()
    stats0 = sim.collision_stats() if sc.collisions else None
    l0 = sim.launch_count()
    t0 = time.perf_counter()
    advance(args.steps)
    sync()
    secs = time.perf_counter() - t0
    line = {"config": config.upper(), "workload": what, "impl": lib.backend, "voxels": sim.n_voxels * world, "links": sim.n_links * world,
            "steps": args.steps, "dt": dt, "ms_per_step": 1e3 * secs / args.steps, "updates_per_s": units * world * args.steps / secs,
            "n_gpus": world if lib.backend.startswith("cuda") else 0,
            "kernel": sim.kernel_name(), "gpu_launches": sim.launch_count() - l0}
    if sc.collisions:
        stats1 = sim.collision_stats()
        f = sim.download("linkflags")
        line.update({"warmup": warm, "collision_pairs_before": stats0[0], "collision_pairs": stats1[0], "rebuilds_before": stats0[1],
                     "rebuilds_in_timed_steps": stats1[1] - stats0[1], "yielded_links": int(((f & 4) != 0).sum()), "failed_links": int(((f & 8) != 0).sum())})
    sim.close()
    return line, secs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c1,c2,c3,c4")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--c3-warmup", type=int, default=8000, help="steps before the timed region of C3 (until contacts are live)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank:
            return
        world = 1
        os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
        lib = capi.load_reference(omp=True)
        sync = lambda: None
        dist = None
    else:
        import torch
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        lib = capi.load_product()

        def sync():
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
                torch.cuda.synchronize()
    for config in args.config.split(","):
        if world > 1 and config != "c4":
            continue                                          # only the ensemble shards without an exchange step
        line, _ = run(lib, config, args, rank, world, sync)
        if rank == 0:
            print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
