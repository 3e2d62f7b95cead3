#!/usr/bin/env python
"""Application-level check at BASELINE.json sizes: the unmodified reference program (oracle/_ref/pagan2_ref) and the
drop-in program (pagan2_msa_b200/_dropin/pagan2_b200: the same program with Viterbi_alignment::align bound to the engine,
one alignment per call) on the same seeded inputs; wall time of each, and every output file compared byte for byte.
    python tools/dropin_walltime.py [--tag r1] [--full]
Sizes: C1 as BASELINE (16 x 1 kb); C4 128 taxa x 300 codons (1000 with --full); C5 8 x 200 kb anchored (32 with --full);
C2 500 reads (placement through the reference's own loop, no batching); C3 100 reads of 400 nt."""
import argparse
import filecmp
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pagan2_msa_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "pagan2_ref")
DROPIN = os.environ.get("PG2_DROPIN_BIN") or os.path.join(ROOT, "pagan2_msa_b200", "_dropin", "pagan2_b200")


def scenarios(full):
    def c1(tmp, rng):
        tree, seqs = synth.balanced_tree(4, synth.random_dna(1000, rng), rng)
        synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
        open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
        return [["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--no-anchors", "--silent"]]

    def c2(tmp, rng):
        tree, seqs = synth.balanced_tree(6, synth.random_dna(1500, rng), rng)
        synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
        open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
        reads = synth.sample_reads(seqs, 500, 150, rng)
        synth.write_fasta(os.path.join(tmp, "r.fas"), reads)
        return [["-s", "s.fas", "-t", "t.nwk", "-o", "ref", "--no-anchors", "--silent"],
                ["--ref-seqfile", "ref.fas", "--ref-treefile", "t.nwk", "--queryfile", "r.fas", "-o", "out", "--no-anchors",
                 "--no-preselection", "--silent"]]

    def c3(tmp, rng):
        t = list(synth.random_dna(2000, rng))
        for i in range(1, len(t)):
            if rng.random() < 0.35:
                t[i] = t[i - 1]
        reads = synth.reads_454("".join(t), 100, 400, rng)
        synth.write_fasta(os.path.join(tmp, "r.fas"), reads)
        return [["--pileup-alignment", "--homopolymer", "--queryfile", "r.fas", "-o", "out", "--no-anchors", "--silent"]]

    def c4(tmp, rng):
        tree, seqs = synth.balanced_codon_tree(7, synth.random_codons(1000 if full else 300, rng), rng)
        synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
        open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
        return [["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--codons", "--no-anchors", "--silent"]]

    def c5(tmp, rng):
        tree, seqs = synth.balanced_tree(5 if full else 3, synth.random_dna(200000, rng), rng, sub=0.01, indel=0.0005)
        synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
        open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
        return [["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--use-prefix-anchors", "--anchors-offset", "15", "--silent"]]

    return [("C1 progressive 16 x 1 kb", c1, 101), ("C2 placement 500 reads x 150 nt, 64-taxon reference", c2, 102),
            ("C3 pileup + homopolymer, 100 reads x 400 nt", c3, 103),
            ("C4 codons 128 taxa x %d codons" % (1000 if full else 300), c4, 104),
            ("C5 anchored %d x 200 kb" % (32 if full else 8), c5, 105)]


def run(binary, fn, seed):
    tmp = tempfile.mkdtemp(prefix="pg2_wall_")
    rng = np.random.default_rng(seed)
    env = dict(os.environ)
    stats = os.path.join(tmp, "b200_stats.json")
    env["PAGAN2_B200_STATS"] = stats
    cmds = fn(tmp, rng)
    times = []
    for args in cmds:
        t0 = time.perf_counter()
        subprocess.run([binary] + args, cwd=tmp, env=env, check=True, timeout=3600, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        times.append(time.perf_counter() - t0)
    outs = sorted(f for f in os.listdir(tmp) if f.startswith("out"))
    st = json.load(open(stats)) if os.path.exists(stats) else None
    return tmp, outs, times[-1], st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r1")
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--only", default="C1", help="substring of the runs to do; '' = all (C2 tests every node for every read and "
                    "C3 / C5 take minutes of host time in the reference program: budget accordingly)")
    args = ap.parse_args()
    rows = []
    for name, fn, seed in scenarios(args.full):
        if args.only and args.only not in name:
            continue
        rd, ro, rt, _ = run(REF, fn, seed)
        nd, no, nt, st = run(DROPIN, fn, seed)
        same = ro == no and len(ro) >= 1 and all(filecmp.cmp(os.path.join(rd, f), os.path.join(nd, f), shallow=False) for f in ro)
        rows.append({"run": name, "reference_wall_s": rt, "dropin_wall_s": nt, "outputs_identical": bool(same), "files": len(ro), "engine": st})
        print(json.dumps(rows[-1]), flush=True)
    with open(os.path.join(ROOT, "gpurun_out" if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else "profiles", args.tag + "_dropin_walltime.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
