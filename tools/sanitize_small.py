#!/usr/bin/env python
"""A small launch batch through every fill kernel and walk, small enough to run under a memory checker where one is
available (compute-sanitizer --tool memcheck python tools/sanitize_small.py; it is closed on the pool this was built on).
Shared-target placement jobs (lane kernel), progressive / pileup / codon jobs (pipelined strips, both step bodies), anchored leaf
pairs (band kernel + segmented walk), the general wavefront kernel; every result is compared with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import enginecheck  # noqa: E402
import randjobs  # noqa: E402
from pagan2_msa_b200 import engine, jobio  # noqa: E402


def main():
    rng = np.random.default_rng(99)
    jobs = []
    for name in ("place_dna", "prog_dna", "pileup_hp", "anchored", "codon"):
        jobs += jobio.load_jobs(os.path.join(ROOT, "tests", "golden", name + ".pjob.gz"))[:3]
    jobs += randjobs.random_shared_target_jobs(rng, 40, nl=60, nr_max=80)
    jobs += [randjobs.random_job(rng, kind) for kind in ("general", "banded", "strip", "banded_chain") for _ in range(4)]
    jobs += [randjobs.random_anchor_band_job(rng, n, bulge=20) for n in (50, 300, 700)]
    jobs = [enginecheck.expect_from_oracle(j) for j in jobs]
    for env in ({}, {"PG2_FORCE_PSRING": "1"}, {"PG2_NO_PSRING": "1", "PG2_NO_BAND": "1"}, {"PG2_FORCE_WAVEFRONT": "1"}):
        os.environ.update(env)
        try:
            with engine.Engine(0) as eng:
                res = enginecheck.check_batch(eng, jobs)
                print(env, "kernels", np.bincount(res["kernel"], minlength=5).tolist(), flush=True)
        finally:
            for k in env:
                os.environ.pop(k, None)


if __name__ == "__main__":
    main()
