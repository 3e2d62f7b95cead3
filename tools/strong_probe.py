#!/usr/bin/env python
"""Tuning aid: what one rank of a strong-scaling run does, on ONE GPU -- the bench workload (--reads alignments in total)
cut for world = 1, 2, 4, 8 by pagan2_msa_b200.shard.partition, rank 0's shard run as a device-resident batch."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from pagan2_msa_b200 import engine, shard  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=100000)
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--targets", action="store_true", help="also the target-aware partitions (shard.partition(groups=...), shard.placement_shards)")
ap.add_argument("--modes", default="", help="comma-separated subset of cells,targets,cost")
ap.add_argument("--worlds", default="1,2,4,8")
args = ap.parse_args()
jobs, info = bench.build_workload(args.reads, 7, 0)
eng = engine.Engine(0)
base = None
for world in [int(x) for x in args.worlds.split(",")]:
    for mode in (args.modes.split(",") if args.modes else (("cells", "targets", "cost") if args.targets else ("cells",))):
        if mode == "cells":
            idx = shard.partition([j.cells for j in jobs], world)[args.rank % world]
        elif mode == "targets":
            idx = shard.partition([j.cells for j in jobs], world, groups=[id(j.left) for j in jobs])[args.rank % world]
        else:
            idx = shard.placement_shards(jobs, world)[args.rank % world]
        mine = [jobs[i] for i in idx]
        cells = sum(j.cells for j in mine)
        b = eng.batch(mine)
        best = None
        for _ in range(4):
            b.run()
            st = eng.stats()
            if best is None or st["run_ms"] < best["run_ms"]:
                best = st
        b.close()
        prep = eng.prepare(mine, pinned=True, compact=True)
        eng.align_prepared(prep)
        e2e = []
        for _ in range(3):
            t0 = time.perf_counter()
            eng.align_prepared(prep)
            e2e.append((time.perf_counter() - t0) * 1e3)
        del prep
        if base is None:
            base = (best["run_ms"], min(e2e))
        print("world %d  partition by %-7s: %6d jobs  %.3g cells  run %.2f ms (fill %.2f, traceback %.2f)  lanes %d (wide shape %d)  "
              "e2e %.2f ms -> strong-scaling efficiency %.2f resident, %.2f e2e" % (
                  world, mode, len(mine), cells, best["run_ms"], best["fill_ms"], best["traceback_ms"], best["jobs_lanes"],
                  best["jobs_lanes_wide"], min(e2e), base[0] / (world * best["run_ms"]), base[1] / (world * min(e2e))), flush=True)
eng.close()
