#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ by running the REFERENCE itself
(oracle/_ref/pagan2_ref = the unmodified reference sources + the job-dump interposer, oracle/Makefile)
on seeded synthetic inputs shaped like BASELINE.json's five configs, at sizes the CPU finishes in
seconds.  Each fixture is a PJOB job stream: every Viterbi_alignment::align call the reference made,
with its inputs (flat CSR graphs, model tables, band) and its outputs (score, full path incl. per-step
scores).  Runs only in the build container (needs /root/reference via oracle/_ref); the fixtures travel.

    python tests/golden/make_golden.py            # regenerate all
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from pagan2_msa_b200 import jobio, synth  # noqa: E402


def thin(jobs, keep, rng):
    if len(jobs) <= keep:
        return jobs
    idx = sorted(rng.choice(len(jobs), size=keep, replace=False).tolist())
    return [jobs[i] for i in idx]


def progressive_dna(tmp):
    """config 1 shape: guide-tree progressive alignment of DNA (-s/-t), 8 taxa x 300 nt."""
    rng = np.random.default_rng(101)
    tree, seqs = synth.balanced_tree(3, synth.random_dna(300, rng), rng)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    jobs, _ = oracle_lib.run_ref(["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--no-anchors", "--silent"], tmp)
    return jobs


def placement_dna(tmp):
    """config 2 shape: reads placed on a reference alignment + tree (trial + final alignments)."""
    rng = np.random.default_rng(202)
    tree, seqs = synth.balanced_tree(3, synth.random_dna(300, rng), rng)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    oracle_lib.run_ref(["-s", "s.fas", "-t", "t.nwk", "-o", "ref", "--no-anchors", "--silent"], tmp, "a.bin", "a.json")
    reads = synth.sample_reads(seqs, 10, 100, rng)
    synth.write_fasta(os.path.join(tmp, "r.fas"), reads)
    jobs, _ = oracle_lib.run_ref(["--ref-seqfile", "ref.fas", "--ref-treefile", "t.nwk", "--queryfile", "r.fas", "-o", "placed",
                                  "--no-anchors", "--test-every-node", "--no-preselection", "--silent"], tmp)
    return thin(jobs, 60, rng)


def pileup_homopolymer(tmp):
    """config 3 shape: --pileup-alignment --homopolymer of 454-like reads (multi-edge read graphs)."""
    rng = np.random.default_rng(303)
    template = synth.random_dna(400, rng)
    # make homopolymer runs common
    t = list(template)
    for i in range(1, len(t)):
        if rng.random() < 0.35:
            t[i] = t[i - 1]
    reads = synth.reads_454("".join(t), 14, 150, rng)
    synth.write_fasta(os.path.join(tmp, "r.fas"), reads)
    jobs, _ = oracle_lib.run_ref(["--pileup-alignment", "--homopolymer", "--queryfile", "r.fas", "-o", "pile", "--no-anchors",
                                  "--silent"], tmp)
    return jobs


def codons(tmp):
    """config 4 shape: --codons progressive alignment (1892-state table)."""
    rng = np.random.default_rng(404)
    tree, seqs = synth.balanced_codon_tree(2, synth.random_codons(60, rng), rng)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    jobs, _ = oracle_lib.run_ref(["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--codons", "--no-anchors", "--silent"], tmp)
    return jobs


def anchored(tmp):
    """config 5 shape: anchored long alignment, prefix anchors -> per-row band."""
    rng = np.random.default_rng(505)
    tree, seqs = synth.balanced_tree(2, synth.random_dna(3000, rng), rng, sub=0.02, indel=0.003)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    jobs, _ = oracle_lib.run_ref(["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--use-prefix-anchors", "--anchors-offset", "15",
                                  "--silent"], tmp)
    return jobs


def bench_targets(tmp):
    """The 127 node graphs (64 leaves + 63 internal nodes of a 64-taxon x 1.5 kb reference alignment,
    built by the reference's Reference_alignment from its own progressive alignment) that bench.py and
    the scale tests place synthetic reads on: the trial alignments of ONE read against every node
    (--test-every-node), node LEFT / read RIGHT (reads_aligner.cpp:3484-3485)."""
    rng = np.random.default_rng(606)
    tree, seqs = synth.balanced_tree(6, synth.random_dna(1500, rng), rng, sub=0.03, indel=0.005)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    oracle_lib.run_ref(["-s", "s.fas", "-t", "t.nwk", "-o", "ref", "--no-anchors", "--silent"], tmp, "a.bin", "a.json", timeout=7200)
    reads = synth.sample_reads(seqs, 1, 150, rng)
    synth.write_fasta(os.path.join(tmp, "r.fas"), reads)
    jobs, _ = oracle_lib.run_ref(["--ref-seqfile", "ref.fas", "--ref-treefile", "t.nwk", "--queryfile", "r.fas", "-o", "placed",
                                  "--no-anchors", "--test-every-node", "--no-preselection", "--silent"], tmp)
    return jobs[:127]


# ---- BASELINE-size job streams of the reference (VERDICT r1, next #1b): stored compactly (jobio, compact=True) ----

def c1_full(tmp):
    """config 1 at BASELINE size: 16 taxa x 1 kb, all 15 alignments (leaf x leaf, ancestor x ancestor up to the root)."""
    rng = np.random.default_rng(1101)
    tree, seqs = synth.balanced_tree(4, synth.random_dna(1000, rng), rng)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    jobs, _ = oracle_lib.run_ref(["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--no-anchors", "--silent"], tmp)
    return jobs


def c3_full(tmp):
    """config 3 at BASELINE read size: --pileup-alignment --homopolymer of 220 454-like 400-nt reads from a 2 kb template;
    the root grows with every read.  Every 10th alignment plus the last four are kept."""
    rng = np.random.default_rng(1303)
    t = list(synth.random_dna(2000, rng))
    for i in range(1, len(t)):
        if rng.random() < 0.35:
            t[i] = t[i - 1]
    reads = synth.reads_454("".join(t), 220, 400, rng)
    synth.write_fasta(os.path.join(tmp, "r.fas"), reads)
    jobs, _ = oracle_lib.run_ref(["--pileup-alignment", "--homopolymer", "--queryfile", "r.fas", "-o", "pile", "--no-anchors",
                                  "--silent"], tmp, timeout=7200)
    keep = sorted(set(range(0, len(jobs), 10)) | set(range(len(jobs) - 4, len(jobs))))
    return [jobs[k] for k in keep]


def c4_full(tmp):
    """config 4 at BASELINE sequence size: --codons, 8 taxa x 1000 codons, all 7 alignments (fas 1892)."""
    rng = np.random.default_rng(1404)
    tree, seqs = synth.balanced_codon_tree(3, synth.random_codons(1000, rng), rng)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    jobs, _ = oracle_lib.run_ref(["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--codons", "--no-anchors", "--silent"], tmp, timeout=7200)
    return jobs


def c5_full(tmp):
    """config 5 at BASELINE sequence size: 4 x 200 kb with prefix anchors: two leaf x leaf and one ancestor x ancestor
    alignment inside their bands, run by the reference itself (~80 s of CPU each)."""
    rng = np.random.default_rng(1505)
    tree, seqs = synth.balanced_tree(2, synth.random_dna(200000, rng), rng, sub=0.01, indel=0.001)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    # the ancestor x ancestor band is wide (the reference predicts more than its default 4 GB of matrices)
    jobs, _ = oracle_lib.run_ref(["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--use-prefix-anchors", "--anchors-offset", "15",
                                  "--memory-for-single-alignment", "50000", "--silent"], tmp, timeout=14400)
    return jobs


FIXTURES = [("prog_dna", progressive_dna), ("place_dna", placement_dna), ("pileup_hp", pileup_homopolymer),
            ("codon", codons), ("anchored", anchored), ("bench_targets", bench_targets),
            ("c1_full", c1_full), ("c3_full", c3_full), ("c4_full", c4_full), ("c5_full", c5_full)]


def prefix_anchor_hits():
    """prefix_anchors.json: the hits of the reference's Find_anchors::find_long_substrings (utils/find_anchors.cpp:35-127) for the
    seeded string pairs of tests/test_anchors.py:cases()."""
    import json

    sys.path.insert(0, os.path.join(HERE, ".."))
    import test_anchors

    out = []
    for a, b, k in test_anchors.cases():
        hits = oracle_lib.ref_prefix_anchors(a, b, k)
        out.append({"len1": len(a), "len2": len(b), "min_length": k, "hits": hits.tolist()})
    with open(os.path.join(HERE, "prefix_anchors.json"), "w") as f:
        json.dump(out, f)
    print("prefix_anchors.json: %d pairs, %d hits" % (len(out), sum(len(o["hits"]) for o in out)))
    # anchor_bands.json: the bands of the reference's Find_anchors::define_tunnel (:320-435) for test_anchors.py:band_cases()
    out = []
    for hits, s1, s2, width in test_anchors.band_cases():
        up, lo = oracle_lib.ref_anchor_band(hits, s1, s2, width)
        out.append({"len1": len(s1), "len2": len(s2), "width": width, "upper": up.tolist(), "lower": lo.tolist()})
    with open(os.path.join(HERE, "anchor_bands.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("anchor_bands.json: %d bands" % len(out))


def main():
    oracle_lib.build_ref()
    only = sys.argv[1:]
    if not only or "prefix_anchors" in only:
        prefix_anchor_hits()
    for name, fn in FIXTURES:
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            jobs = fn(tmp)
        path = os.path.join(HERE, name + ".pjob.gz")
        jobio.save_jobs(path, jobs, compact=name.endswith("_full"))
        cells = sum(j.cells for j in jobs)
        banded = sum(1 for j in jobs if j.upper is not None)
        print("%-10s %4d jobs %10d cells %3d banded  fas=%d  %8.1f KB" % (name, len(jobs), cells, banded, jobs[0].model.fas,
                                                                       os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
