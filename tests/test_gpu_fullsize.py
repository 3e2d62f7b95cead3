"""GPU parity at BASELINE.json's FULL sizes for the configs that are not the bench line (the bench covers configs[1]):
each job is aligned through the C-ABI and compared bit for bit -- score, every path field, every per-step score --
with the oracle (oracle/viterbi_oracle.c, itself pinned to the reference on the golden dumps)."""
import numpy as np
import pytest

import enginecheck
import randjobs
from pagan2_msa_b200 import abi, engine, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__

    __graft_entry__.build()
    e = engine.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_old():
    """Without the pipelined-strip and band kernels (strip + wavefront kernels only)."""
    import os

    os.environ["PG2_NO_PSTRIP"] = "1"
    os.environ["PG2_NO_BAND"] = "1"
    try:
        e = engine.Engine(0)
    finally:
        os.environ.pop("PG2_NO_PSTRIP", None)
        os.environ.pop("PG2_NO_BAND", None)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_ps():
    """Banded chain x chain jobs on the pipelined-strip kernel too (by default they stay on the wavefront kernel)."""
    import os

    os.environ["PG2_PSTRIP_BANDED_CHAINS"] = "1"
    try:
        e = engine.Engine(0)
    finally:
        os.environ.pop("PG2_PSTRIP_BANDED_CHAINS", None)
    yield e
    e.close()


def indel_copy(a, rng, sub, n_indels, max_len):
    """b = a with substitutions and a few indels; returns (b, col) with col[i] = column of b facing row i of a."""
    pos = np.sort(rng.choice(np.arange(100, len(a) - 100), size=n_indels, replace=False))
    out, col, j, prev = [], np.zeros(len(a), np.int64), 0, 0
    for p in pos:
        seg = a[prev:p].copy()
        out.append(seg)
        col[prev:p] = j + np.arange(p - prev)
        j += p - prev
        n = int(rng.integers(1, max_len + 1))
        if rng.random() < 0.5:  # insertion in b
            out.append(rng.integers(0, 4, size=n).astype(a.dtype))
            j += n
            prev = p
        else:  # deletion from b
            col[p:p + n] = j
            prev = p + n
    out.append(a[prev:].copy())
    col[prev:] = j + np.arange(len(a) - prev)
    b = np.concatenate(out)
    mut = rng.random(len(b)) < sub
    b[mut] = rng.integers(0, 4, size=int(mut.sum()))
    return b, col


def test_config5_anchored_200kb(eng, eng_old, eng_ps, golden):
    """configs[4]: one 200 kb x 200 kb alignment inside an anchor band (10 M in-band cells): the band kernel, the pipelined
    strips, and the wavefront kernel's chain path."""
    rng = np.random.default_rng(5005)
    model = golden["anchored"][0].model
    a = rng.integers(0, 4, size=200000).astype(np.int32)
    b, col = indel_copy(a, rng, 0.02, 40, 12)
    left, right = abi.FlatGraph.chain(a), abi.FlatGraph.chain(b)
    lx, ly = left.n_sites - 1, right.n_sites - 1
    c = np.concatenate([[0], col + 1])[:lx]  # DP row i faces DP column col[i-1]+1
    job = abi.FlatJob(left, right, model, 2)
    job.upper = (c - 25).astype(np.int32)
    job.lower = (c + 25).astype(np.int32)
    job.lower[-1] = max(job.lower[-1], ly + 3)
    job.upper = np.maximum.accumulate(job.upper).astype(np.int32)
    job.lower = np.maximum.accumulate(job.lower).astype(np.int32)
    assert 9_000_000 < job.cells < 12_000_000
    job = enginecheck.expect_from_oracle(job)
    res = enginecheck.check_batch(eng, [job])
    assert res["kernel"][0] == 4 and res["status"][0] == 0
    res = enginecheck.check_batch(eng_ps, [job])
    assert res["kernel"][0] == 3 and res["status"][0] == 0
    res = enginecheck.check_batch(eng_old, [job])
    assert res["kernel"][0] == 0 and res["status"][0] == 0


def test_config4_codons_1000(eng, eng_old, golden):
    """configs[3]: 1002 x 1002 codon sites, 1892-state table (strip kernel, global float table), leaf vs leaf and
    an ancestor-shaped graph vs leaf."""
    rng = np.random.default_rng(4004)
    model = golden["codon"][0].model
    assert model.fas == 1892
    a = rng.integers(0, 61, size=1000).astype(np.int32)
    b = a.copy()
    mut = rng.random(1000) < 0.1
    b[mut] = rng.integers(0, 61, size=int(mut.sum()))
    b = np.delete(b, rng.choice(1000, size=12, replace=False))
    anc = randjobs.random_graph(rng, 1000, 61, p_extra=0.05)
    jobs = [abi.FlatJob(abi.FlatGraph.chain(a), abi.FlatGraph.chain(b), model, 2),
            abi.FlatJob(anc, abi.FlatGraph.chain(b), model, 2)]
    jobs = [enginecheck.expect_from_oracle(j) for j in jobs]
    res = enginecheck.check_batch(eng, jobs)
    assert (res["kernel"] == 3).all()
    res = enginecheck.check_batch(eng_old, jobs)
    assert (res["kernel"] == 1).all()


def test_config3_pileup_shape(eng, eng_old, golden):
    """configs[2]: a grown pileup root (2 k sites, multi-edge) against a 400-nt 454 read graph (multi-edge): the
    general wavefront kernel, both graphs with long-span edges and weights."""
    rng = np.random.default_rng(3003)
    model = golden["pileup_hp"][0].model
    jobs = []
    for _ in range(3):
        left = randjobs.random_graph(rng, 2000, 4, p_extra=0.15, max_span=6)
        right = randjobs.random_graph(rng, 400, 4, p_extra=0.3, max_span=4)
        jobs.append(abi.FlatJob(left, right, model, 2))
    jobs = [enginecheck.expect_from_oracle(j) for j in jobs]
    res = enginecheck.check_batch(eng, jobs)
    assert np.isin(res["kernel"], (0, 3)).all()
    res = enginecheck.check_batch(eng_old, jobs)
    assert (res["kernel"] == 0).all()


def test_config1_progressive_1kb_wave(eng, eng_old, golden):
    """configs[0]: the 8 + 4 + 2 + 1 alignments of a 16-taxon x 1 kb guide tree as launch batches: leaves (plain
    chains) and ancestor-shaped graphs on both sides."""
    rng = np.random.default_rng(1001)
    model = golden["prog_dna"][0].model
    root = synth.random_dna(1000, rng)
    leaves = [synth.dna_states(synth.evolve(root, rng)) for _ in range(16)]
    wave1 = [abi.FlatJob(abi.FlatGraph.chain(leaves[2 * k]), abi.FlatGraph.chain(leaves[2 * k + 1]), model, 2) for k in range(8)]
    wave1 = [enginecheck.expect_from_oracle(j) for j in wave1]
    res = enginecheck.check_batch(eng, wave1)
    assert (res["kernel"] == 3).all()
    res = enginecheck.check_batch(eng_old, wave1)
    assert (res["kernel"] == 1).all()
    anc = [randjobs.random_graph(rng, 1100, 4, p_extra=0.06, max_span=8) for _ in range(6)]
    wave2 = [abi.FlatJob(anc[2 * k], anc[2 * k + 1], model, 2) for k in range(3)]
    wave2 = [enginecheck.expect_from_oracle(j) for j in wave2]
    res = enginecheck.check_batch(eng, wave2)
    assert (res["kernel"] == 3).all()
    res = enginecheck.check_batch(eng_old, wave2)
    assert (res["kernel"] == 0).all()


def test_wavefront_diagonals_longer_than_the_ring(eng_old, golden):
    eng = eng_old
    """A general x general job whose anti-diagonals (3000 cells) exceed the wavefront kernel's shared-memory ring
    (WAVE_RING_MAX = 2816): every score read goes to the global scratch, the original path."""
    rng = np.random.default_rng(2816)
    model = golden["prog_dna"][0].model
    left = randjobs.random_graph(rng, 3000, 4, p_extra=0.05, max_span=8)
    right = randjobs.random_graph(rng, 3000, 4, p_extra=0.05, max_span=8)
    res = enginecheck.check_batch(eng, [enginecheck.expect_from_oracle(abi.FlatJob(left, right, model, 2))])
    assert res["kernel"][0] == 0 and res["status"][0] == 0
