"""End-to-end drop-in check: the reference command-line program with Viterbi_alignment::align bound to the engine
(pagan2_msa_b200/host, link-time interposer) must write byte-identical output files to the unmodified reference binary
(oracle/_ref/pagan2_ref) on the same seeded inputs -- progressive DNA, anchored, pileup + homopolymer, codon and
query placement runs.

CPU variant: the host mirror linked against the test emulation of the engine (tests/_emu/pagan2_b200_emu) -- checks
the host mirror (graph packing, path expansion, used-edge marks, build_ancestral_sequence hand-over).
GPU variant (-m gpu): the real binary pagan2_msa_b200/_dropin/pagan2_b200 over libpagan2_b200.so.
Both binaries are prebuilt in the build container (they need /root/reference) and travel to the GPU box."""
import filecmp
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib
from pagan2_msa_b200 import abi, synth

REF = os.path.join(abi.REPO_ROOT, "oracle", "_ref", "pagan2_ref")
DROPIN = os.path.join(abi.REPO_ROOT, "pagan2_msa_b200", "_dropin", "pagan2_b200")
DROPIN_EMU = os.path.join(abi.REPO_ROOT, "tests", "_emu", "pagan2_b200_emu")


def build_binaries():
    if os.path.isdir("/root/reference/src"):
        oracle_lib.build_ref()
        subprocess.check_call(["make", "-s", "-C", os.path.join(abi.REPO_ROOT, "pagan2_msa_b200", "csrc")],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(abi.REPO_ROOT, "pagan2_msa_b200", "host")])
        subprocess.check_call(["make", "-s", "-C", os.path.join(abi.REPO_ROOT, "tests", "emu"), "../_emu/pagan2_b200_emu"])


def scenario_progressive(tmp, rng):
    tree, seqs = synth.balanced_tree(3, synth.random_dna(250, rng), rng)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    return [["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--no-anchors", "--silent"]]


def scenario_anchored(tmp, rng):
    tree, seqs = synth.balanced_tree(2, synth.random_dna(1500, rng), rng, sub=0.02, indel=0.003)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    return [["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--use-prefix-anchors", "--anchors-offset", "15", "--silent"]]


def scenario_pileup(tmp, rng):
    t = list(synth.random_dna(300, rng))
    for i in range(1, len(t)):
        if rng.random() < 0.35:
            t[i] = t[i - 1]
    reads = synth.reads_454("".join(t), 10, 120, rng)
    synth.write_fasta(os.path.join(tmp, "r.fas"), reads)
    return [["--pileup-alignment", "--homopolymer", "--queryfile", "r.fas", "-o", "out", "--no-anchors", "--silent"]]


def scenario_codons(tmp, rng):
    tree, seqs = synth.balanced_codon_tree(2, synth.random_codons(40, rng), rng)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    return [["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--codons", "--no-anchors", "--silent"]]


def scenario_placement(tmp, rng):
    tree, seqs = synth.balanced_tree(3, synth.random_dna(250, rng), rng)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    reads = synth.sample_reads(seqs, 8, 90, rng)
    synth.write_fasta(os.path.join(tmp, "r.fas"), reads)
    return [["-s", "s.fas", "-t", "t.nwk", "-o", "ref", "--no-anchors", "--silent"],
            ["--ref-seqfile", "ref.fas", "--ref-treefile", "t.nwk", "--queryfile", "r.fas", "-o", "out", "--no-anchors",
             "--test-every-node", "--no-preselection", "--silent"]]


def scenario_placement_fragments(tmp, rng):
    """--fragments: every read is first mapped to its node (static trial alignments: reads x candidate nodes), then the
    reads of a node are aligned to it (reads_aligner.cpp:372-621)."""
    tree, seqs = synth.balanced_tree(3, synth.random_dna(250, rng), rng)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    reads = synth.sample_reads(seqs, 12, 90, rng)
    synth.write_fasta(os.path.join(tmp, "r.fas"), reads)
    return [["-s", "s.fas", "-t", "t.nwk", "-o", "ref", "--no-anchors", "--silent"],
            ["--ref-seqfile", "ref.fas", "--ref-treefile", "t.nwk", "--queryfile", "r.fas", "-o", "out", "--no-anchors",
             "--test-every-node", "--no-preselection", "--fragments", "--silent"]]


def scenario_progressive16(tmp, rng):
    tree, seqs = synth.balanced_tree(4, synth.random_dna(200, rng), rng)
    synth.write_fasta(os.path.join(tmp, "s.fas"), seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    return [["-s", "s.fas", "-t", "t.nwk", "-o", "out", "--no-anchors", "--silent"]]


SCENARIOS = {"placement_fragments": (scenario_placement_fragments, 16), "progressive16": (scenario_progressive16, 17),
             "progressive": (scenario_progressive, 11), "anchored": (scenario_anchored, 12), "pileup": (scenario_pileup, 13),
             "codons": (scenario_codons, 14), "placement": (scenario_placement, 15)}


def run_program(binary, name, extra_args=(), extra_env=None):
    fn, seed = SCENARIOS[name]
    tmp = tempfile.mkdtemp(prefix="pg2_e2e_")
    rng = np.random.default_rng(seed)
    env = dict(os.environ)
    env.update(extra_env or {})
    stats = os.path.join(tmp, "b200_stats.json")
    env["PAGAN2_B200_STATS"] = stats
    for args in fn(tmp, rng):
        subprocess.run([binary] + args + list(extra_args), cwd=tmp, env=env, check=True, timeout=1800, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
    outs = sorted(f for f in os.listdir(tmp) if f.startswith("out") or f.startswith("ref."))
    return tmp, outs, stats


def compare(binary, name, extra_args=(), extra_env=None, ref_args=()):
    """The reference program (its default, serial traversal unless ref_args says otherwise) against `binary` run with
    extra_args / extra_env: every output file byte for byte."""
    ref_dir, ref_outs, _ = run_program(REF, name, ref_args)
    new_dir, new_outs, stats = run_program(binary, name, extra_args, extra_env)
    assert ref_outs == new_outs and len(ref_outs) >= 1, (ref_outs, new_outs)
    for f in ref_outs:
        assert filecmp.cmp(os.path.join(ref_dir, f), os.path.join(new_dir, f), shallow=False), "%s differs (%s)" % (f, name)
    import json

    with open(stats) as fh:
        st = json.load(fh)
    assert st["jobs"] >= 1 and st["cells"] > 0  # the alignments really went through the engine
    return st


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_dropin_matches_reference_cpu_emulation(name):
    build_binaries()
    if not (os.path.exists(REF) and os.path.exists(DROPIN_EMU)):
        pytest.skip("reference binaries are built in the container that has /root/reference")
    compare(DROPIN_EMU, name)


def check_schedulers(binary):
    """The batch schedulers of the host mirror against the reference's serial runs, byte for byte:
    --threads N sends the guide tree through the wave scheduler (one launch batch per wave of ready nodes), placement
    prefetches the trial alignments of the coming reads x candidate nodes in one launch batch, and PAGAN2_B200_DEVICES
    cuts every launch batch over several contexts."""
    st = compare(binary, "progressive16", extra_args=["--threads", "4"])
    assert st["jobs"] == 15 and st["wave_batches"] == 3 and st["batches"] == 4, st  # waves of 8, 4, 2 and the root
    assert st["model_cache_hits"] == 14, st  # one branch length everywhere: alignment_model runs once, 14 deep copies
    st2 = compare(binary, "progressive16", extra_env={"PAGAN2_B200_NO_MODEL_CACHE": "1"})
    assert st2["model_cache_hits"] == 0 and st2["batches"] == 15, st2  # (and the serial traversal: one alignment per batch)
    st = compare(binary, "progressive16", extra_args=["--threads", "3", "--boost"])
    assert st["jobs"] == 15 and st["batches"] == 4, st
    st = compare(binary, "anchored", extra_args=["--threads", "2"])
    assert st["wave_batches"] >= 1, st
    st = compare(binary, "codons", extra_args=["--threads", "2"])
    assert st["wave_batches"] >= 1, st
    for name in ("placement_fragments", "placement"):
        st = compare(binary, name)
        assert st["prefetch_batches"] >= 1 and st["cache_hits"] > st["prefetch_batches"] and st["batches"] < st["jobs"], st
        st2 = compare(binary, name, extra_env={"PAGAN2_B200_NO_PREFETCH": "1"})
        assert st2["prefetch_batches"] == 0 and st2["cache_hits"] == 0
    st = compare(binary, "placement_fragments", extra_env={"PAGAN2_B200_PREFETCH_READS": "5"})  # several windows
    assert st["prefetch_batches"] >= 3, st
    st = compare(binary, "progressive16", extra_args=["--threads", "4"], extra_env={"PAGAN2_B200_DEVICES": "0,0"})
    assert st["sharded_batches"] >= 2, st  # the waves of 8 and 4 (a batch of two stays on one device)
    st = compare(binary, "placement_fragments", extra_env={"PAGAN2_B200_DEVICES": "0,0,0"})
    assert st["sharded_batches"] >= 1 and st["cache_hits"] > 0, st


def test_dropin_schedulers_cpu_emulation():
    build_binaries()
    if not (os.path.exists(REF) and os.path.exists(DROPIN_EMU)):
        pytest.skip("reference binaries are built in the container that has /root/reference")
    check_schedulers(DROPIN_EMU)


@pytest.mark.gpu
def test_dropin_schedulers_on_b200():
    assert os.path.exists(REF) and os.path.exists(DROPIN), "prebuilt binaries missing: run __graft_entry__.build() where /root/reference exists"
    check_schedulers(DROPIN)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_dropin_matches_reference_on_b200(name):
    assert os.path.exists(REF) and os.path.exists(DROPIN), "prebuilt binaries missing: run __graft_entry__.build() where /root/reference exists"
    compare(DROPIN, name)
