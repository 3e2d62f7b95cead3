#!/usr/bin/env python
"""Tuning aid: wall time of pg2_align_batch on the bench workload (--timing prints the host packing steps and the device
timeline of the chunks); --sweep tries several chunk cuts of the pipelined call in one process."""
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


def run(label, reps=4):
    best = 1e9
    for rep in range(reps):
        t0 = time.perf_counter(); eng.align_prepared(prep); t1 = time.perf_counter()
        best = min(best, (t1 - t0) * 1e3)
        st = eng.stats()
        print("%s pg2_align_batch (pinned result buffers): %.1f ms   h2d %.1f ms  kernels %.1f ms  d2h %.1f ms" % (label, (t1 - t0) * 1e3, st["h2d_ms"], st["run_ms"], st["d2h_ms"]), flush=True)
    return best


if "--sweep" in sys.argv:
    out = {}
    for w in ("", "1,1", "1,3,3", "1,2,4,4,4", "1,2,3,3,3,3,3,2", "1,2,3,3,3,3,3,3,3,3,2,1", "1,2,2,2,2,2,2,2,2,2,2,2,2,2,2,1"):
        if w:
            os.environ["PG2_PIPELINE_WEIGHTS"] = w
        else:
            os.environ.pop("PG2_PIPELINE_WEIGHTS", None)
        out[w or "default"] = run("[weights %s]" % (w or "default"), 3)
    print(out)
else:
    run("")
