#!/usr/bin/env python
"""Tuning aid: wall time of the three parts of pg2_align_batch (create = host packing, run = upload + kernels,
fetch = device->host + result unpack) on the bench workload."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from pagan2_msa_b200 import engine
jobs, info = bench.build_workload(100000, 7)
eng = engine.Engine(0)
for rep in range(3):
    t0 = time.perf_counter(); b = eng.batch(jobs); t1 = time.perf_counter()
    b.run(); eng.lib.pg2_stream_synchronize(eng.ctx); t2 = time.perf_counter()
    st = eng.stats()
    res, steps = b.fetch(); t3 = time.perf_counter()
    b.close()
    print("create %.1f ms (python struct build incl.)  run %.1f ms (h2d %.1f, kernels %.1f)  fetch %.1f ms (d2h %.1f)" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, st["h2d_ms"], st["run_ms"], (t3 - t2) * 1e3, eng.stats()["d2h_ms"]))
prep = eng.prepare(jobs, pinned=True)
for rep in range(3):
    t0 = time.perf_counter(); eng.align_prepared(prep); t1 = time.perf_counter()
    print("pg2_align_batch (pinned result buffers): %.1f ms" % ((t1 - t0) * 1e3))
