"""Seeded synthetic workloads shaped like BASELINE.json's configs.

Pure generators of sequences / trees / reads and of flat jobs built from them.  No alignment logic:
the graphs built here are the reference's *leaf* graph shapes (Sequence::create_default_sequence,
sequence.cpp:152-303), which are fully determined by the input string.
"""
import numpy as np

from .abi import FlatGraph, FlatJob

DNA = "ACGT"
DNA_FULL_ALPHABET = "ACGTRYMKWSBDHVN"  # Model_factory::dna_full_char_alphabet order (model_factory.cpp:102-103)

STOP_CODONS = {"TAA", "TAG", "TGA"}


def random_dna(n, rng):
    return "".join(DNA[i] for i in rng.integers(0, 4, size=n))


def evolve(seq, rng, sub=0.05, indel=0.01, ext=0.7):
    """One branch: substitutions with prob `sub` per site, geometric indels opened with prob `indel`."""
    out = []
    i = 0
    n = len(seq)
    while i < n:
        r = rng.random()
        if r < indel / 2:
            l = 1
            while rng.random() < ext:
                l += 1
            i += l
            continue
        if r < indel:
            l = 1
            while rng.random() < ext:
                l += 1
            out.extend(DNA[k] for k in rng.integers(0, 4, size=l))
        c = seq[i]
        if rng.random() < sub:
            c = DNA[rng.integers(0, 4)]
        out.append(c)
        i += 1
    return "".join(out)


def balanced_tree(depth, root_seq, rng, sub=0.05, indel=0.01, branch=0.05, prefix="t"):
    """Balanced binary tree of 2**depth taxa; returns (newick, [(name, seq)])."""

    def build(seq, d, name):
        if d == 0:
            return name, [(name, seq)]
        l, ls = build(evolve(seq, rng, sub, indel), d - 1, name + "a")
        r, rs = build(evolve(seq, rng, sub, indel), d - 1, name + "b")
        return "(%s:%g,%s:%g)" % (l, branch, r, branch), ls + rs

    t, seqs = build(root_seq, depth, prefix)
    return t + ";", seqs


def random_codons(n, rng):
    out = []
    while len(out) < n:
        c = random_dna(3, rng)
        if c not in STOP_CODONS:
            out.append(c)
    return "".join(out)


def evolve_codons(seq, rng, sub=0.05, indel=0.01):
    cods = [seq[i : i + 3] for i in range(0, len(seq), 3)]
    out = []
    i = 0
    while i < len(cods):
        r = rng.random()
        if r < indel / 2:
            i += 1 + int(rng.integers(0, 3))
            continue
        if r < indel:
            for _ in range(1 + int(rng.integers(0, 3))):
                out.append(random_codons(1, rng))
        c = cods[i]
        if rng.random() < sub * 3:
            for _ in range(10):
                p = int(rng.integers(0, 3))
                c2 = c[:p] + DNA[rng.integers(0, 4)] + c[p + 1 :]
                if c2 not in STOP_CODONS:
                    c = c2
                    break
        out.append(c)
        i += 1
    return "".join(out)


def balanced_codon_tree(depth, root_seq, rng, sub=0.03, indel=0.01, branch=0.05, prefix="c"):
    def build(seq, d, name):
        if d == 0:
            return name, [(name, seq)]
        l, ls = build(evolve_codons(seq, rng, sub, indel), d - 1, name + "a")
        r, rs = build(evolve_codons(seq, rng, sub, indel), d - 1, name + "b")
        return "(%s:%g,%s:%g)" % (l, branch, r, branch), ls + rs

    t, seqs = build(root_seq, depth, prefix)
    return t + ";", seqs


def sample_reads(seqs, n_reads, read_len, rng, sub=0.01):
    """150-nt style reads: substrings of the leaf sequences with `sub` substitutions."""
    reads = []
    for k in range(n_reads):
        name, s = seqs[int(rng.integers(0, len(seqs)))]
        st = int(rng.integers(0, max(1, len(s) - read_len)))
        r = list(s[st : st + read_len])
        for q in range(len(r)):
            if rng.random() < sub:
                r[q] = DNA[rng.integers(0, 4)]
        reads.append(("read%d" % k, "".join(r), name))
    return reads


def reads_454(template, n_reads, read_len, rng, hp_err=0.15, sub=0.005):
    """454-like reads: homopolymer-length errors (+-1) with prob hp_err per run, rare substitutions."""
    reads = []
    for k in range(n_reads):
        st = int(rng.integers(0, max(1, len(template) - read_len)))
        s = template[st : st + read_len]
        out = []
        i = 0
        while i < len(s):
            j = i
            while j < len(s) and s[j] == s[i]:
                j += 1
            run = j - i
            if run >= 2 and rng.random() < hp_err:
                run += 1 if rng.random() < 0.5 else -1
            c = s[i]
            if rng.random() < sub:
                c = DNA[rng.integers(0, 4)]
            out.append(c * max(run, 1))
            i = j
        reads.append(("hp%d" % k, "".join(out)))
    return reads


def write_fasta(path, entries):
    with open(path, "w") as f:
        for e in entries:
            f.write(">%s\n%s\n" % (e[0], e[1]))


def dna_states(seq):
    return np.array([DNA_FULL_ALPHABET.index(c) for c in seq], dtype=np.int32)


def leaf_graph(seq):
    """Plain DNA leaf graph of the reference (no 454/homopolymer edges)."""
    return FlatGraph.chain(dna_states(seq))


def placement_jobs(targets, reads_states, assignment, model, flags=2):
    """Query-placement launch batch: job k aligns target graph assignment[k] (LEFT, as in
    Reads_aligner::create_temp_node, reads_aligner.h:169-184) with read k (RIGHT).
    `targets` are FlatGraph objects that are shared between jobs (uploaded once)."""
    jobs = []
    for k, st in enumerate(reads_states):
        jobs.append(FlatJob(targets[assignment[k]], FlatGraph.chain(st), model, flags))
    return jobs
