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
        reads = synth.sample_reads(seqs, 2000 if full else 200, 150, rng)
        synth.write_fasta(os.path.join(tmp, "r.fas"), reads)
        return [["-s", "s.fas", "-t", "t.nwk", "-o", "ref", "--no-anchors", "--silent"],
                ["--ref-seqfile", "ref.fas", "--ref-treefile", "t.nwk", "--queryfile", "r.fas", "-o", "out", "--no-anchors",
                 "--no-preselection", "--test-every-terminal-node", "--fragments", "--silent"]]

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
        # (the reference predicts more than its default 4 GB of matrices for the ancestor x ancestor bands)
        return [["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--use-prefix-anchors", "--anchors-offset", "15",
                 "--memory-for-single-alignment", "50000", "--silent"]]

    return [("C1 progressive 16 x 1 kb", c1, 101), ("C2 placement %d reads x 150 nt, 64-taxon reference, every leaf tested" % (2000 if full else 200), c2, 102),
            ("C3 pileup + homopolymer, 100 reads x 400 nt", c3, 103),
            ("C4 codons 128 taxa x %d codons" % (1000 if full else 300), c4, 104),
            ("C5 anchored %d x 200 kb" % (32 if full else 8), c5, 105)]


def sha256_file(path):
    import hashlib

    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def run(binary, fn, seed, extra_args=(), extra_env=None):
    tmp = tempfile.mkdtemp(prefix="pg2_wall_")
    rng = np.random.default_rng(seed)
    env = dict(os.environ)
    env.update(extra_env or {})
    stats = os.path.join(tmp, "b200_stats.json")
    env["PAGAN2_B200_STATS"] = stats
    cmds = fn(tmp, rng)
    times = []
    for k, args in enumerate(cmds):
        t0 = time.perf_counter()
        # the scheduler flags go to the run being measured (the last command; the first of a placement run builds the reference alignment)
        extra = list(extra_args) if k == len(cmds) - 1 else []
        subprocess.run([binary] + args + extra, cwd=tmp, env=env, check=True, timeout=14400, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        times.append(time.perf_counter() - t0)
    outs = sorted(f for f in os.listdir(tmp) if f.startswith("out"))
    st = json.load(open(stats)) if os.path.exists(stats) else None
    return tmp, outs, times[-1], st


EXPECTED = os.path.join(ROOT, "tests", "golden", "dropin_expected.json")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r2")
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--only", default="C1", help="substring of the runs to do; '' = all")
    ap.add_argument("--reference-only", action="store_true",
                    help="run only the reference program and record the SHA-256 of every output file and its wall time in "
                         "tests/golden/dropin_expected.json (minutes to hours of CPU: done once, in the build container)")
    ap.add_argument("--dropin-only", action="store_true",
                    help="run only the drop-in program and compare its outputs with the recorded SHA-256s")
    ap.add_argument("--dropin-args", default="", help="extra arguments for the drop-in program, e.g. '--threads 8'")
    ap.add_argument("--devices", default="", help="PAGAN2_B200_DEVICES for the drop-in program")
    args = ap.parse_args()
    expected = json.load(open(EXPECTED)) if os.path.exists(EXPECTED) else {}
    rows = []
    for name, fn, seed in scenarios(args.full):
        if args.only and args.only not in name:
            continue
        key = name + (" [full]" if args.full else "")
        if args.reference_only:
            rd, ro, rt, _ = run(REF, fn, seed)
            expected[key] = {"reference_wall_s": rt, "host": "%d cores (build container)" % os.cpu_count(),
                             "files": {f: sha256_file(os.path.join(rd, f)) for f in ro}}
            with open(EXPECTED, "w") as f:
                json.dump(expected, f, indent=1, sort_keys=True)
            print(json.dumps({"run": key, **expected[key]}), flush=True)
            continue
        env = {"PAGAN2_B200_DEVICES": args.devices} if args.devices else None
        if args.dropin_only:
            if key not in expected:
                print(json.dumps({"run": key, "skipped": "no recorded reference outputs"}), flush=True)
                continue
            nd, no, nt, st = run(DROPIN, fn, seed, args.dropin_args.split(), env)
            want = expected[key]["files"]
            same = sorted(want) == no and all(sha256_file(os.path.join(nd, f)) == want[f] for f in no)
            rows.append({"run": key, "reference_wall_s": expected[key]["reference_wall_s"], "reference_host": expected[key]["host"],
                         "dropin_wall_s": nt, "dropin_args": args.dropin_args, "outputs_identical": bool(same), "files": len(no), "engine": st})
        else:
            rd, ro, rt, _ = run(REF, fn, seed)
            nd, no, nt, st = run(DROPIN, fn, seed, args.dropin_args.split(), env)
            same = ro == no and len(ro) >= 1 and all(filecmp.cmp(os.path.join(rd, f), os.path.join(nd, f), shallow=False) for f in ro)
            rows.append({"run": key, "reference_wall_s": rt, "dropin_wall_s": nt, "dropin_args": args.dropin_args,
                         "outputs_identical": bool(same), "files": len(ro), "engine": st})
        print(json.dumps(rows[-1]), flush=True)
    if rows:
        out_dir = os.path.join(ROOT, "gpurun_out" if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else "profiles")
        with open(os.path.join(out_dir, args.tag + "_dropin_walltime.json"), "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
