#!/usr/bin/env python
"""Tuning aid: device-resident fill time of the bench workload for several builds of the CUDA library.
    python tools/fillbench.py [--reads N] lib_a.so lib_b.so ...   (default: the product library)"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from pagan2_msa_b200 import engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=100000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("libs", nargs="*")
args = ap.parse_args()
jobs, info = bench.build_workload(args.reads, 7)
for lib in args.libs or [engine.LIB_PATH]:
    eng = engine.Engine(0, lib)
    b = eng.batch(jobs)
    best = None
    for _ in range(args.reps):
        b.run()
        st = eng.stats()
        if best is None or st["fill_ms"] < best["fill_ms"]:
            best = st
    res, _ = b.fetch()
    b.close()
    eng.close()
    print("%-40s fill %.2f ms  traceback %.2f ms  run %.2f ms  -> fill %.1f GCUPS  ok=%d lanes=%d strip=%d" % (
        os.path.basename(lib), best["fill_ms"], best["traceback_ms"], best["run_ms"], info["cells_per_step"] / best["fill_ms"] * 1e-6,
        int((res["status"] == 0).sum()), best["jobs_lanes"], best["jobs_strip"]), flush=True)
