"""Prefix anchors: pg2_find_prefix_anchors against the reference's own Find_anchors::find_long_substrings
(utils/find_anchors.cpp:35-127, run live through oracle/_ref) and against committed hit lists of the reference."""
import json
import os

import numpy as np
import pytest

import oracle_lib
from pagan2_msa_b200 import abi, engine, synth

GOLDEN = os.path.join(abi.REPO_ROOT, "tests", "golden", "prefix_anchors.json")


def related_pair(rng, n, sub, indel, alphabet=b"ACGT"):
    a = bytes(rng.choice(list(alphabet), size=n).astype(np.uint8))
    out = bytearray()
    for c in a:
        u = rng.random()
        if u < indel:
            continue
        if u < 2 * indel:
            out.append(int(rng.choice(list(alphabet))))
        out.append(int(rng.choice(list(alphabet))) if rng.random() < sub else c)
    return a, bytes(out)


def cases():
    rng = np.random.default_rng(2024)
    out = []
    for n, sub, indel, k in ((300, 0.02, 0.002, 12), (2000, 0.03, 0.001, 20), (5000, 0.01, 0.0005, 30), (20000, 0.03, 0.0005, 30),
                             (1200, 0.2, 0.01, 8), (900, 0.0, 0.0, 30)):
        a, b = related_pair(rng, n, sub, indel)
        out.append((a, b, k))
    a, b = related_pair(rng, 3000, 0.05, 0.002, alphabet=b"ACDEFGHIKLMNPQRSTVWY")
    out.append((a, b, 10))
    out.append((b"ACGTACGTAC", b"TTTTTTTT", 4))          # no hit
    out.append((b"ACGTTGCAAGGCT", b"ACGTTGCAAGGCT", 5))  # identical: one hit, and tied suffixes
    out.append((b"", b"ACGT", 2))
    out.append((b"AAAAAAAAAAAAAAAAAAAA", b"AAAAAAAAAAAAAAAAAAAAAAAA", 3))  # low complexity
    return out


def test_matches_committed_reference_hits():
    """(the fixture was written by tests/golden/make_golden.py --anchors from the live reference)"""
    want = json.load(open(GOLDEN))
    assert len(want) == len(cases())
    for (a, b, k), w in zip(cases(), want):
        got = engine.find_prefix_anchors(a, b, k)
        assert got.tolist() == w["hits"], (len(a), len(b), k)


@pytest.mark.skipif(not oracle_lib.ref_available(), reason="oracle/_ref not built")
def test_matches_live_reference():
    for a, b, k in cases():
        got = engine.find_prefix_anchors(a, b, k)
        ref = oracle_lib.ref_prefix_anchors(a, b, k)
        assert got.tolist() == ref.tolist(), (len(a), len(b), k)


@pytest.mark.skipif(not oracle_lib.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(6))
def test_matches_live_reference_all_length_residues(seed):
    """Lengths of every residue modulo 16 (the reference keeps its text in stack arrays rounded up to 16 bytes and writes the
    terminator one past the end, find_anchors.cpp:43-61)."""
    rng = np.random.default_rng(100 + seed)
    for n in range(640 + seed, 640 + seed + 96, 6):
        a, b = related_pair(rng, n, 0.03, 0.003)
        b = b[: len(b) - int(rng.integers(0, 16))]
        got = engine.find_prefix_anchors(a, b, 12)
        ref = oracle_lib.ref_prefix_anchors(a, b, 12)
        assert got.tolist() == ref.tolist(), (len(a), len(b))


@pytest.mark.skipif(not oracle_lib.ref_available(), reason="oracle/_ref not built")
def test_200kb_pair_matches_live_reference_and_is_fast():
    import time

    rng = np.random.default_rng(7)
    a, b = related_pair(rng, 200000, 0.01, 0.0005)
    t0 = time.perf_counter()
    got = engine.find_prefix_anchors(a, b, 30)
    t1 = time.perf_counter()
    ref = oracle_lib.ref_prefix_anchors(a, b, 30)
    t2 = time.perf_counter()
    assert got.tolist() == ref.tolist()
    assert len(got) > 500
    print("200 kb pair: %d hits, %.2f s here, %.2f s in the reference" % (len(got), t1 - t0, t2 - t1))
    assert (t1 - t0) < (t2 - t1)


def gapped(rng, s, p):
    """Sequence string with the '-' characters of skipped sites (Sequence::get_sequence_string(true))."""
    out = bytearray()
    for c in s:
        while rng.random() < p:
            out.append(ord("-"))
        out.append(c)
    return bytes(out)


def band_cases():
    rng = np.random.default_rng(77)
    out = []
    for n, sub, indel, k, width, pgap in ((400, 0.03, 0.004, 12, 15, 0.0), (3000, 0.02, 0.001, 20, 15, 0.02), (3000, 0.02, 0.001, 20, 5, 0.0),
                                          (8000, 0.05, 0.002, 30, 15, 0.01), (200, 0.3, 0.02, 30, 15, 0.0), (50, 0.0, 0.0, 10, 80, 0.0),
                                          (1000, 0.01, 0.0, 25, 0, 0.0)):
        a, b = related_pair(rng, n, sub, indel)
        hits = engine.find_prefix_anchors(a, b, k)
        # what define_tunnel gets: the hits that survive check_hits_order_conflict are a subset in start order; a sorted copy
        # and the raw list are both valid inputs of the function
        out.append((hits, gapped(rng, a, pgap), gapped(rng, b, pgap), width))
        out.append((hits[np.argsort(hits[:, 0], kind="stable")] if len(hits) else hits, gapped(rng, a, pgap), gapped(rng, b, pgap), width))
    return out


@pytest.mark.skipif(not oracle_lib.ref_available(), reason="oracle/_ref not built")
def test_band_matches_live_reference():
    for hits, s1, s2, width in band_cases():
        up, lo = engine.anchor_band(hits, s1, s2, width)
        rup, rlo = oracle_lib.ref_anchor_band(hits, s1, s2, width)
        assert up.tolist() == rup.tolist() and lo.tolist() == rlo.tolist(), (len(s1), len(s2), width, len(hits))


def test_band_matches_committed_reference_bands():
    want = json.load(open(GOLDEN.replace("prefix_anchors", "anchor_bands")))
    assert len(want) == len(band_cases())
    for (hits, s1, s2, width), w in zip(band_cases(), want):
        up, lo = engine.anchor_band(hits, s1, s2, width)
        assert up.tolist() == w["upper"] and lo.tolist() == w["lower"]


@pytest.mark.skipif(not oracle_lib.ref_available(), reason="oracle/_ref not built")
def test_band_200kb_matches_live_reference_and_is_fast():
    import time

    rng = np.random.default_rng(8)
    a, b = related_pair(rng, 200000, 0.01, 0.0005)
    hits = engine.find_prefix_anchors(a, b, 30)
    hits = hits[np.argsort(hits[:, 0], kind="stable")]
    t0 = time.perf_counter()
    up, lo = engine.anchor_band(hits, a, b, 15)
    t1 = time.perf_counter()
    rup, rlo = oracle_lib.ref_anchor_band(hits, a, b, 15)
    t2 = time.perf_counter()
    assert up.tolist() == rup.tolist() and lo.tolist() == rlo.tolist()
    print("200 kb band: %.3f s here, %.2f s in the reference" % (t1 - t0, t2 - t1))
    assert (t1 - t0) < (t2 - t1)
