"""CPU tests of the oracle: pinned against the reference's own outputs (golden dumps; and, where
oracle/_ref is present, the reference library run live on random flat jobs)."""
import numpy as np
import pytest

import oracle_lib
import randjobs


def check_job(job, ref_used=None):
    status, score, steps, cells, ul, ur = oracle_lib.oracle_align(job, marks=True)
    assert status == 0
    assert cells == job.cells
    assert np.float64(score).view(np.uint64) == np.float64(job.expected_score).view(np.uint64)
    assert oracle_lib.steps_equal(steps, job.expected_path, job.expected_path_score, getattr(job, 'expected_path_score_sha', None)) == []
    # is_used edge marks (viterbi_alignment.cpp:1054-1155): the restatement's marks, as a set, are what the reference
    # itself left marked -- on fresh graphs (live reference) or as the difference the dump recorded around its call
    if ref_used is not None:
        assert sorted(set(ul.tolist())) == ref_used[0].tolist() and sorted(set(ur.tolist())) == ref_used[1].tolist()
    used = getattr(job, "expected_used", None)
    if used is not None:
        for side, mine in (("l", ul), ("r", ur)):
            before, after = set(used[side][0].tolist()), set(used[side][1].tolist())
            assert before | set(mine.tolist()) == after


@pytest.mark.parametrize("name", ["prog_dna", "place_dna", "pileup_hp", "codon", "anchored"])
def test_oracle_matches_reference_dump(golden, name):
    for job in golden[name]:
        check_job(job)


@pytest.mark.parametrize("name", ["c1_full", "c3_full", "c4_full", "c5_full"])
def test_oracle_matches_reference_at_baseline_size(golden, name):
    """The restatement against job streams the REFERENCE ran at BASELINE sizes (16 x 1 kb, 400-nt pileup, 1000 codons,
    200 kb anchored incl. ancestor x ancestor inside its band -- where FP64 rounding first matters)."""
    for job in golden[name]:
        check_job(job)


def test_golden_covers_all_shapes(golden):
    assert any(j.upper is not None for j in golden["anchored"])
    assert golden["codon"][0].model.fas == 1892
    deg = lambda g: int(np.diff(g.off).max())
    assert max(deg(j.right) for j in golden["pileup_hp"]) >= 2      # homopolymer multi-edge reads
    assert max(deg(j.left) for j in golden["place_dna"]) >= 2       # internal reference nodes
    assert any((j.left.logw != 0).any() for j in golden["place_dna"])


@pytest.mark.skipif(not oracle_lib.ref_available(), reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("kind", ["general", "banded", "strip"])
def test_oracle_matches_live_reference_on_random_jobs(kind):
    rng = np.random.default_rng({"general": 11, "banded": 12, "strip": 13}[kind])
    for _ in range(40):
        job = randjobs.random_job(rng, kind)
        score, path, pscore = oracle_lib.ref_align_flat(job)
        job.expected_score, job.expected_path, job.expected_path_score = score, path, pscore
        check_job(job, oracle_lib.ref_last_used())


@pytest.mark.skipif(not oracle_lib.ref_available(), reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("shape", ["banded_chain", "shared_target"])
def test_oracle_matches_live_reference_on_more_shapes(shape):
    """Further pins of the restatement against the reference library run live: plain chains inside a band (the anchored
    leaf x leaf shape), reads against one shared ancestor-shaped target (the placement shape) and parameter sets made of a few repeated values (ties everywhere: the first-wins order is what is being checked)."""
    rng = np.random.default_rng({"banded_chain": 31, "shared_target": 33}[shape])
    jobs = []
    if shape == "banded_chain":
        jobs = [randjobs.random_job(rng, "banded_chain") for _ in range(30)]
    else:
        for _ in range(3):
            jobs += randjobs.random_shared_target_jobs(rng, 12)
    for job in jobs:
        score, path, pscore = oracle_lib.ref_align_flat(job)
        job.expected_score, job.expected_path, job.expected_path_score = score, path, pscore
        check_job(job, oracle_lib.ref_last_used())


def test_oracle_rejects_bad_band():
    rng = np.random.default_rng(5)
    job = randjobs.random_job(rng, "banded")
    job.upper = job.upper.copy()
    job.upper[len(job.upper) // 2] = job.upper[-1] + 5  # not monotone
    status, _, _, _ = oracle_lib.oracle_align(job)
    assert status == 2


def test_oracle_no_path_in_disconnected_band():
    rng = np.random.default_rng(6)
    job = randjobs.random_job(rng, "general")
    lx, ly = job.left.n_sites - 1, job.right.n_sites - 1
    if lx < 4 or ly < 4:
        job = randjobs.random_job(rng, "general")
        lx, ly = job.left.n_sites - 1, job.right.n_sites - 1
    job.upper = np.zeros(lx, np.int32)
    job.lower = np.zeros(lx, np.int32)  # only column 0: the end corner is unreachable unless ly == 1
    status, score, _, _ = oracle_lib.oracle_align(job)
    assert (status == 1 and score == -np.inf) or ly == 1
