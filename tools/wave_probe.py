#!/usr/bin/env python
"""Tuning aid: one launch batch of general x general jobs (C1 wave 2 shape) on the wavefront kernel, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import configs_report as cr
from pagan2_msa_b200 import engine
which = sys.argv[1] if len(sys.argv) > 1 else "C1 wave 2"
name, jobs = [b for b in cr.batches() if b[0].startswith(which)][0]
eng = engine.Engine(0)
b = eng.batch(jobs)
for _ in range(3):
    b.run()
    st = eng.stats()
    print(name, "fill %.2f ms traceback %.2f ms" % (st["fill_ms"], st["traceback_ms"]), flush=True)
b.close(); eng.close()
