#!/usr/bin/env python
"""Device time, GCUPS and wall time per alignment for the launch batches of the five BASELINE.json configs at their
full sizes (shapes as in tests/test_gpu_fullsize.py; configs[1] is bench.py's own line), next to the C oracle on one
host core for one job of each batch.  Writes profiles/<tag>_configs.json and prints a markdown table.
    python tools/configs_report.py --tag r1"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
import randjobs  # noqa: E402
import test_gpu_fullsize as fs  # noqa: E402
from pagan2_msa_b200 import abi, engine, jobio, synth  # noqa: E402

KERNEL = {0: "wavefront", 1: "strip", 2: "lanes", 3: "pstrip", 4: "band"}


def golden(name):
    return jobio.load_jobs(os.path.join(ROOT, "tests", "golden", name + ".pjob.gz"))


def batches():
    rng = np.random.default_rng(11)
    out = []
    # C1: 16 taxa x 1 kb -- wave 1: 8 leaf pairs; waves 2-4: ancestor x ancestor (the reference's own 1.5 kb ancestor graphs)
    m = golden("prog_dna")[0].model
    root = synth.random_dna(1000, rng)
    leaves = [synth.dna_states(synth.evolve(root, rng)) for _ in range(16)]
    out.append(("C1 wave 1: 8 x (1 kb leaf x 1 kb leaf)",
                [abi.FlatJob(abi.FlatGraph.chain(leaves[2 * k]), abi.FlatGraph.chain(leaves[2 * k + 1]), m, 2) for k in range(8)]))
    tg = [j.left for j in golden("bench_targets")]
    anc = [g for g in tg if not ((np.diff(g.off)[1:] == 1).all() and (g.logw == 0).all())]
    out.append(("C1 wave 2: 4 x (ancestor x ancestor, 1.5 k sites)", [abi.FlatJob(anc[2 * k], anc[2 * k + 1], m, 2) for k in range(4)]))
    # C3: pileup + homopolymer: sequential, one alignment per launch (grown root x 454 read graph)
    m3 = golden("pileup_hp")[0].model
    out.append(("C3 pileup: 1 x (2 k-site root x 400-nt 454 read), both multi-edge",
                [abi.FlatJob(randjobs.random_graph(rng, 2000, 4, p_extra=0.15, max_span=6),
                             randjobs.random_graph(rng, 400, 4, p_extra=0.3, max_span=4), m3, 2)]))
    # C4: codons, wave 1 of 128 taxa: 64 leaf pairs of 1000 codons, 1892-state table
    m4 = golden("codon")[0].model
    jobs = []
    for _ in range(64):
        a = rng.integers(0, 61, size=1000).astype(np.int32)
        b = a.copy()
        mut = rng.random(1000) < 0.1
        b[mut] = rng.integers(0, 61, size=int(mut.sum()))
        b = np.delete(b, rng.choice(1000, size=12, replace=False))
        jobs.append(abi.FlatJob(abi.FlatGraph.chain(a), abi.FlatGraph.chain(b), m4, 2))
    out.append(("C4 wave 1: 64 x (1000 codons x 1000 codons), fas 1892", jobs))
    out.append(("C4 wave 2: 32 x (ancestor-shaped x ancestor-shaped, 1000 codons)",
                [abi.FlatJob(randjobs.random_graph(rng, 1000, 61, p_extra=0.05), randjobs.random_graph(rng, 1000, 61, p_extra=0.05), m4, 2)
                 for _ in range(32)]))
    # C5: anchored 200 kb, wave 1 of 32 sequences: 16 banded alignments
    m5 = golden("anchored")[0].model
    jobs = []
    for _ in range(16):
        a = rng.integers(0, 4, size=200000).astype(np.int32)
        b, col = fs.indel_copy(a, rng, 0.02, 40, 12)
        left, right = abi.FlatGraph.chain(a), abi.FlatGraph.chain(b)
        lx, ly = left.n_sites - 1, right.n_sites - 1
        c = np.concatenate([[0], col + 1])[:lx]
        job = abi.FlatJob(left, right, m5, 2)
        job.upper = (c - 25).astype(np.int32)
        job.lower = (c + 25).astype(np.int32)
        job.lower[-1] = max(job.lower[-1], ly + 3)
        job.upper = np.maximum.accumulate(job.upper).astype(np.int32)
        job.lower = np.maximum.accumulate(job.lower).astype(np.int32)
        jobs.append(job)
    out.append(("C5 wave 1: 16 x (200 kb x 200 kb inside a +-25 anchor band)", jobs))
    # the reference's own job streams at BASELINE sizes (tests/golden/*_full: dumped from the reference program)
    c1 = golden("c1_full")
    depth = {}
    for k, j in enumerate(c1):  # a node's wave = how many of its two children are ancestors, by graph shape: leaves are plain chains
        plain = lambda g: bool((np.diff(g.off)[1:] == 1).all())
        depth[k] = 0 if plain(j.left) and plain(j.right) else 1
    out.append(("C1 (reference stream) wave 1: 8 leaf x leaf, 1 kb", [j for k, j in enumerate(c1) if depth[k] == 0]))
    out.append(("C1 (reference stream) waves 2-4: 7 ancestor x ancestor", [j for k, j in enumerate(c1) if depth[k] == 1]))
    out.append(("C1 (reference stream) root: 1 alignment, 1.2 k x 1.2 k sites", [c1[-1]]))
    c3 = golden("c3_full")
    out.append(("C3 (reference stream): 1 x (root after 220 reads x 400-nt 454 read)", [c3[-1]]))
    c4 = golden("c4_full")
    out.append(("C4 (reference stream): 4 leaf x leaf, 1000 codons", c4[:1] + c4[1:2] + [j for j in c4[2:] if (np.diff(j.left.off)[1:] == 1).all()][:2]))
    out.append(("C4 (reference stream) root: 1 ancestor x ancestor, 1000 codons", [c4[-1]]))
    c5 = golden("c5_full")
    out.append(("C5 (reference stream): 2 x (200 kb leaf x leaf, reference's anchor band)", c5[:2]))
    out.append(("C5 (reference stream): 1 x (200 kb ancestor x ancestor, reference's anchor band)", c5[2:]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r1")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="", help="substring filter on the batch name")
    ap.add_argument("--max-cells", type=float, default=0, help="skip batches with more cells (the wavefront kernel needs 36 B/cell)")
    args = ap.parse_args()
    eng = engine.Engine(0)
    rows = []
    for name, jobs in batches():
        if args.only and args.only not in name:
            continue
        if args.max_cells and sum(j.cells for j in jobs) > args.max_cells:
            continue
        cells = int(sum(j.cells for j in jobs))
        b = eng.batch(jobs)
        best = None
        for _ in range(args.reps + 1):
            b.run()
            st = eng.stats()
            if best is None or st["run_ms"] < best["run_ms"]:
                best = st
        res, _ = b.fetch()
        b.close()
        prep = eng.prepare(jobs, pinned=True)
        eng.align_prepared(prep)
        t = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            eng.align_prepared(prep)
            t.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        status, score, _, _ = oracle_lib.oracle_align(jobs[0])
        cpu_s = time.perf_counter() - t0
        ok = bool(np.float64(res["score"][0]).view(np.uint64) == np.float64(score).view(np.uint64))
        kinds = sorted(set(KERNEL[int(k)] for k in res["kernel"]))
        rows.append({"batch": name, "jobs": len(jobs), "cells": cells, "kernel": "+".join(kinds),
                     "fill_ms": best["fill_ms"], "traceback_ms": best["traceback_ms"], "device_ms": best["run_ms"],
                     "device_gcups": cells / best["run_ms"] * 1e-6, "e2e_ms": min(t) * 1e3, "e2e_gcups": cells / min(t) * 1e-9,
                     "e2e_ms_per_alignment": min(t) * 1e3 / len(jobs),
                     "cpu_oracle_1core_s_per_alignment": cpu_s, "cpu_oracle_mcups": jobs[0].cells / cpu_s * 1e-6,
                     "first_job_score_matches_oracle": ok})
        print(json.dumps(rows[-1]), flush=True)
    eng.close()
    with open(os.path.join(ROOT, "profiles", args.tag + "_configs.json"), "w") as f:
        json.dump(rows, f, indent=1)
    print("\n| launch batch | jobs | cells | kernel | device ms | device GCUPS | e2e ms | e2e ms / alignment | CPU oracle, 1 core: s / alignment (MCUPS) |")
    print("|---|---:|---:|---|---:|---:|---:|---:|---:|")
    for r in rows:
        print("| %s | %d | %.3g | %s | %.2f | %.2f | %.2f | %.3f | %.2f (%.1f) |" % (
            r["batch"], r["jobs"], r["cells"], r["kernel"], r["device_ms"], r["device_gcups"], r["e2e_ms"], r["e2e_ms_per_alignment"],
            r["cpu_oracle_1core_s_per_alignment"], r["cpu_oracle_mcups"]))


if __name__ == "__main__":
    main()
