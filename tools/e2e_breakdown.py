#!/usr/bin/env python
"""Tuning aid: wall time of the parts of pg2_align_batch on the bench workload (--timing prints the host packing steps)."""
import os, sys, time
if "--timing" in sys.argv:
    os.environ["PG2_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from pagan2_msa_b200 import engine
jobs, info = bench.build_workload(100000, 7)
eng = engine.Engine(0)
prep = eng.prepare(jobs, pinned=True, compact="--explicit" not in sys.argv)
for rep in range(4):
    t0 = time.perf_counter(); eng.align_prepared(prep); t1 = time.perf_counter()
    st = eng.stats()
    print("pg2_align_batch (pinned result buffers): %.1f ms   h2d %.1f ms  kernels %.1f ms  d2h %.1f ms" % ((t1 - t0) * 1e3, st["h2d_ms"], st["run_ms"], st["d2h_ms"]), flush=True)
