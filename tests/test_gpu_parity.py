"""GPU parity tests (run with -m gpu on a B200): the CUDA path through the C-ABI against the committed
golden dumps of the reference, against the oracle on seeded random jobs, and -- at BASELINE config sizes --
through size-independent properties."""
import os

import numpy as np
import pytest

import enginecheck
import oracle_lib
import randjobs
from pagan2_msa_b200 import abi, engine, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__

    __graft_entry__.build()
    e = engine.Engine(0)  # raises if the CUDA library or the device is missing: no fallback
    yield e
    e.close()


def engine_with(**env):
    """An Engine created under the given PG2_* switches (read once, at pg2_ctx_create)."""
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        return engine.Engine(0)
    finally:
        for k in env:
            os.environ.pop(k, None)


@pytest.fixture(scope="module")
def eng_wave():
    e = engine_with(PG2_FORCE_WAVEFRONT=1)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_old():
    """Without the pipelined-strip kernel: the warp-per-alignment strip kernel and the general wavefront kernel (the
    fallback for graphs outside the pipelined-strip kernel's limits) stay covered."""
    e = engine_with(PG2_NO_PSTRIP=1, PG2_NO_BAND=1)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_ps():
    """Every job that is not a lane task goes to the pipelined-strip kernel."""
    e = engine_with(PG2_NO_LANES=1, PG2_PSTRIP_MAX_JOBS=1000000, PG2_PSTRIP_BANDED_CHAINS=1)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_ps4():
    e = engine_with(PG2_NO_LANES=1, PG2_PSTRIP_MAX_JOBS=1000000, PG2_PSTRIP_K=4, PG2_PSTRIP_BANDED_CHAINS=1)
    yield e
    e.close()


@pytest.mark.parametrize("name", ["prog_dna", "place_dna", "pileup_hp", "codon", "anchored"])
def test_golden_default_dispatch(eng, golden, name):
    enginecheck.check_batch(eng, golden[name])


@pytest.mark.parametrize("name", ["prog_dna", "place_dna", "pileup_hp", "codon", "anchored"])
def test_golden_wavefront_kernel(eng_wave, golden, name):
    res = enginecheck.check_batch(eng_wave, golden[name])
    assert (res["kernel"] == 0).all()


@pytest.mark.parametrize("name", ["prog_dna", "place_dna", "pileup_hp", "codon", "anchored"])
def test_golden_pstrip_kernel(eng_ps, eng_ps4, eng_old, golden, name):
    """The reference's job streams of all five config shapes through the pipelined-strip kernel (both strip widths) and
    through the older kernels only."""
    for e in (eng_ps, eng_ps4):
        res = enginecheck.check_batch(e, golden[name])
        assert (res["kernel"] == 3).all()
    res = enginecheck.check_batch(eng_old, golden[name])
    assert (res["kernel"] != 3).all()


@pytest.mark.parametrize("name", ["c1_full", "c3_full", "c4_full", "c5_full"])
def test_reference_streams_at_baseline_size(eng, eng_old, golden, name):
    """BASELINE-size job streams dumped from the REFERENCE itself (tests/golden/make_golden.py): 16 x 1 kb progressive (all
    15 alignments), the growing pileup root against 400-nt 454 read graphs, 1000-codon alignments with the 1892-state
    table, and 200 kb x 200 kb anchored alignments incl. ancestor x ancestor inside its band.  Score bits, every path
    field, every per-step score (or their SHA-256 for the 400 000-step paths) and the used-edge marks."""
    jobs = golden[name]
    res = enginecheck.check_batch(eng, jobs)
    assert (res["kernel"] == 3).sum() >= len(jobs) - 2
    if name != "c5_full":  # the wavefront kernel needs 36 bytes per cell: minutes for the widest 200 kb band
        enginecheck.check_batch(eng_old, jobs)


@pytest.fixture(scope="module")
def eng_nolanes():
    e = engine_with(PG2_NO_LANES=1, PG2_NO_PSTRIP=1)
    yield e
    e.close()


def test_placement_uses_register_strip_kernels(eng, eng_nolanes, golden):
    res = enginecheck.check_batch(eng, golden["place_dna"])
    assert np.isin(res["kernel"], (1, 2, 3)).all()  # (fewer than LANE_MIN_JOBS reads per target here: no lane tasks)
    res = enginecheck.check_batch(eng_nolanes, golden["place_dna"])
    assert (res["kernel"] == 1).all()


@pytest.mark.parametrize("seed,plain_left,n_jobs", [(161, False, 70), (162, True, 40), (163, False, 33), (164, False, 16), (165, True, 100)])
@pytest.mark.parametrize("lane_w", [4, 10])  # three CTAs of 4 warps per SM / one CTA of 10: the throughput and the latency shape
def test_lane_kernel_shared_target_vs_oracle(eng, monkeypatch, lane_w, seed, plain_left, n_jobs):
    """Jobs sharing the left graph run one alignment per lane (pg2_lanes.cu): all template variants, ragged read
    lengths, thin remainders, singletons left to the strip kernel."""
    monkeypatch.setenv("PG2_LANE_W", str(lane_w))
    rng = np.random.default_rng(seed)
    jobs = []
    for _ in range(6):
        jobs += randjobs.random_shared_target_jobs(rng, n_jobs, plain_left=plain_left)
    jobs += [randjobs.random_job(rng, "strip") for _ in range(5)]
    jobs = [enginecheck.expect_from_oracle(j) for j in jobs]
    res = enginecheck.check_batch(eng, jobs)
    assert (res["kernel"][:6 * n_jobs] == 2).all()
    assert np.isin(res["kernel"][-5:], (1, 3)).all()
    st = eng.stats()
    assert st["jobs_lanes_wide"] == (st["jobs_lanes"] if lane_w == 10 else 0)


@pytest.mark.parametrize("lane_w", [4, 10])
def test_lane_kernel_long_reads_and_bad_job(eng, monkeypatch, lane_w):
    monkeypatch.setenv("PG2_LANE_W", str(lane_w))
    rng = np.random.default_rng(166)
    jobs = randjobs.random_shared_target_jobs(rng, 40, nl=60, nr_max=200)
    jobs = [enginecheck.expect_from_oracle(j) for j in jobs]
    jobs[7].right.state[1] = 99
    jobs[7].expected_status = abi.PG2_JOB_BAD_GRAPH
    res = enginecheck.check_batch(eng, jobs)
    assert (res["kernel"] == 2).all()


def test_lane_and_strip_kernels_agree_bitwise(eng, eng_nolanes, golden):
    """The same placement batch through the lane kernel and through the strip kernel: identical bits."""
    model = golden["place_dna"][0].model
    rng = np.random.default_rng(167)
    targets = [j.left for j in golden["place_dna"][:4]]
    reads, assign = [], []
    for k in range(400):
        a = int(rng.integers(0, len(targets)))
        seq = targets[a].state[1:-1]
        st = int(rng.integers(0, max(1, len(seq) - 150)))
        r = seq[st:st + int(rng.integers(30, 151))].copy()
        r = np.where(r < 4, r, 0).astype(np.int32)
        reads.append(r)
        assign.append(a)
    jobs = synth.placement_jobs(targets, reads, assign, model)
    ra, sa = eng.align(jobs)
    rb, sb = eng_nolanes.align(jobs)
    assert (ra["kernel"] == 2).all() and (rb["kernel"] == 1).all()
    assert (ra["status"] == 0).all()
    assert (ra["score"].view(np.uint64) == rb["score"].view(np.uint64)).all()
    for k, job in enumerate(jobs):
        pa, _, _ = eng.expand(job, ra[k], sa)
        pb, _, _ = eng_nolanes.expand(job, rb[k], sb)
        assert pa.tobytes() == pb.tobytes()


@pytest.mark.parametrize("kind,seed", [("general", 121), ("banded", 122), ("strip", 123), ("banded_chain", 124)])
def test_random_jobs_vs_oracle(eng, eng_old, eng_ps4, kind, seed):
    rng = np.random.default_rng(seed)
    jobs = [enginecheck.expect_from_oracle(randjobs.random_job(rng, kind)) for _ in range(300)]
    res = enginecheck.check_batch(eng, jobs)
    if kind == "banded_chain":
        assert (res["kernel"] == 4).all()  # plain unit-weight chains inside a band: the band kernel
    else:
        assert (res["kernel"] == 3).mean() > 0.9
    res = enginecheck.check_batch(eng_ps4, jobs)
    assert (res["kernel"] == 3).mean() > 0.9
    res = enginecheck.check_batch(eng_old, jobs)
    assert (res["kernel"] != 3).all()


@pytest.mark.parametrize("seed,banded", [(191, False), (192, True), (193, False), (194, True)])
def test_pstrip_many_blocks(eng, eng_ps4, seed, banded):
    """Several column blocks per alignment pipelined over the warps of a CTA: general x general graphs of a few hundred
    to 1500 sites, long-span edges across lanes, parked rows, bands that leave some blocks without rows."""
    rng = np.random.default_rng(seed)
    jobs = []
    for n_max in (420, 420, 420, 1500, 1500, 3000):
        model = randjobs.random_model(rng, 15, ties=rng.random() < 0.4)
        nl, nr = int(rng.integers(150, n_max)), int(rng.integers(150, n_max))
        left = randjobs.random_graph(rng, nl, 15, p_extra=0.1, max_span=int(rng.integers(3, 25)))
        right = randjobs.random_graph(rng, nr, 15, p_extra=float(rng.choice([0.03, 0.12])), max_span=int(rng.integers(3, 20)))
        job = abi.FlatJob(left, right, model, int(rng.integers(0, 4)))
        if banded:
            job.upper, job.lower = randjobs.random_band(rng, left.n_sites - 1, right.n_sites - 1, min_w=3, max_w=40)
        jobs.append(enginecheck.expect_from_oracle(job))
    for e in (eng, eng_ps4):
        res = enginecheck.check_batch(e, jobs)
        assert (res["kernel"] == 3).sum() >= len(jobs) - 1


def test_both_kernels_agree_on_strip_jobs(eng, eng_wave):
    rng = np.random.default_rng(77)
    jobs = [randjobs.random_job(rng, "strip") for _ in range(200)]
    ra, sa = eng.align(jobs)
    rb, sb = eng_wave.align(jobs)
    assert np.isin(ra["kernel"], (1, 2, 3)).all() and (rb["kernel"] == 0).all()
    assert (ra["score"].view(np.uint64) == rb["score"].view(np.uint64)).all()
    for k, job in enumerate(jobs):
        pa, _, _ = eng.expand(job, ra[k], sa)
        pb, _, _ = eng_wave.expand(job, rb[k], sb)
        assert pa.tobytes() == pb.tobytes()


def test_bad_inputs_statuses(eng):
    rng = np.random.default_rng(31)
    good = enginecheck.expect_from_oracle(randjobs.random_job(rng, "general"))
    bad_graph = randjobs.random_job(rng, "general")
    bad_graph.left.start[-1] = bad_graph.left.n_sites + 5
    bad_graph.expected_status = abi.PG2_JOB_BAD_GRAPH
    bad_band = randjobs.random_job(rng, "banded")
    bad_band.upper = bad_band.upper.copy()
    bad_band.upper[len(bad_band.upper) // 2] = bad_band.upper[-1] + 5
    bad_band.expected_status = abi.PG2_JOB_BAD_BAND
    enginecheck.check_batch(eng, [good, bad_graph, bad_band, good])


def test_placement_config_scale_properties(eng, golden):
    """BASELINE configs[1] shape at scale: 4096 reads x 150 nt against 1.5 kb targets.
    Properties that need no oracle: (1) a read cut from the target aligns with an all-match path and
    the path is a monotone walk ending at the corner; (2) identical jobs give identical bits;
    (3) the replayed path score equals the device score (checked inside pg2_expand_path);
    (4) a sample is compared with the oracle bit for bit."""
    model = golden["place_dna"][0].model
    rng = np.random.default_rng(2024)
    targets = [synth.leaf_graph(synth.random_dna(1500, rng)) for _ in range(8)]
    tseqs = [t.state[1:-1] for t in targets]
    reads, assign = [], []
    for k in range(4096):
        a = int(rng.integers(0, 8))
        st = int(rng.integers(0, 1350))
        r = tseqs[a][st:st + 150].copy()
        if k % 2:
            mut = rng.random(150) < 0.02
            r[mut] = rng.integers(0, 4, size=int(mut.sum()))
        reads.append(r)
        assign.append(a)
    jobs = synth.placement_jobs(targets, reads, assign, model)
    res, steps = eng.align(jobs)
    assert (res["status"] == 0).all() and (res["kernel"] == 2).all()
    res2, steps2 = eng.align(jobs)
    assert res.tobytes() == res2.tobytes() and steps.tobytes() == steps2.tobytes()
    for k in range(0, 4096, 64):
        p, ul, ur = eng.expand(jobs[k], res[k], steps)  # raises if the replayed score differs by one bit
        real = p[p["real_site"] == 1]
        assert (np.diff(real["x_ind"]) >= 0).all() and (np.diff(real["y_ind"]) >= 0).all()
        assert (real["matrix"] == abi.PG2_M_MAT).sum() >= 140  # the read matches its source window
        if k % 2 == 0:
            m = real[real["matrix"] == abi.PG2_M_MAT]
            assert len(m) == 150 and (np.diff(m["x_ind"]) == 1).all()
    sample = [enginecheck.expect_from_oracle(jobs[k]) for k in range(0, 4096, 256)]
    enginecheck.check_batch(eng, sample)


def test_progressive_config_scale_vs_oracle(eng, eng_old, eng_ps):
    """BASELINE configs[0] shape: 1 kb x 1 kb leaf alignments (multi-block strip) and a banded
    200 kb-style corridor job on the band kernel (and on the wavefront kernel), both against the oracle."""
    rng = np.random.default_rng(9)
    model = randjobs.random_model(rng, 15)
    a = synth.random_dna(1000, rng)
    b = synth.evolve(a, rng)
    job = abi.FlatJob(synth.leaf_graph(a), synth.leaf_graph(b), model, 2)
    lx, ly = job.left.n_sites - 1, job.right.n_sites - 1
    banded = abi.FlatJob(synth.leaf_graph(a), synth.leaf_graph(b), model, 2)
    banded.upper, banded.lower = randjobs.random_band(rng, lx, ly, 20, 40)
    jobs = [enginecheck.expect_from_oracle(job), enginecheck.expect_from_oracle(banded)]
    res = enginecheck.check_batch(eng, jobs)
    assert res["kernel"][0] == 3 and res["kernel"][1] == 4  # plain unit-weight chains inside a band: the band kernel
    res = enginecheck.check_batch(eng_ps, jobs)
    assert (res["kernel"] == 3).all()
    res = enginecheck.check_batch(eng_old, jobs)
    assert res["kernel"][0] == 1 and res["kernel"][1] == 0


def test_single_job_batches_on_a_fresh_engine(golden):
    """A batch of ONE strip-kernel job on a fresh context: the launch rounds up to a whole CTA, and every warp of
    it must own scratch (the drop-in binary aligns one job per call)."""
    for k in (0, 7, 13):
        with engine.Engine(0) as e:
            enginecheck.check_batch(e, [golden["place_dna"][k]])
            enginecheck.check_batch(e, [golden["pileup_hp"][min(k, len(golden["pileup_hp"]) - 1)]])


def test_pipelined_align_batch_on_device(eng):
    """The chunked, double-buffered pg2_align_batch path (two contexts, two streams) against the oracle."""
    rng = np.random.default_rng(181)
    jobs = []
    for _ in range(6):
        jobs += randjobs.random_shared_target_jobs(rng, 70)
    jobs += [randjobs.random_job(rng, kind) for kind in ("strip", "general", "banded") for _ in range(10)]
    rng.shuffle(jobs)
    jobs = [enginecheck.expect_from_oracle(j) for j in jobs]
    os.environ["PG2_PIPELINE_MIN_JOBS"] = "8"
    os.environ["PG2_PIPELINE_CHUNKS"] = "4"
    try:
        res = enginecheck.check_batch(eng, jobs)
        assert (res["kernel"] == 2).sum() >= 6 * 64
        enginecheck.check_batch(eng, jobs[::-1])
    finally:
        os.environ.pop("PG2_PIPELINE_MIN_JOBS", None)
        os.environ.pop("PG2_PIPELINE_CHUNKS", None)


def test_compact_chain_form_on_device(eng, golden):
    """Plain chains in the compact pg2_graph form (states only) against the explicit form and the oracle."""
    rng = np.random.default_rng(191)
    jobs = list(golden["place_dna"][:30]) + randjobs.random_shared_target_jobs(rng, 70, weights=False)
    jobs += [randjobs.random_job(rng, kind) for kind in ("banded_chain", "strip") for _ in range(10)]
    jobs = [enginecheck.expect_from_oracle(j) for j in jobs]
    ra = enginecheck.check_batch(eng, jobs).copy()
    rb, sb = eng.align_prepared(eng.prepare(jobs, compact=True))
    assert ra.tobytes() == rb.tobytes()
    for k, job in enumerate(jobs):
        if rb["status"][k] == 0:
            st, _, _ = eng.expand(job, rb[k], sb, compact=True)
            assert oracle_lib.steps_equal(st, job.expected_path, job.expected_path_score) == []


def test_lane_shapes_agree_at_scale(eng, monkeypatch):
    """A shard of the bench workload (the 127 nodes of the reference tree as targets, root-most ancestors with a third of
    their sites multi-edge) through the lane kernel's throughput shape and its latency shape: every score, status and
    encoded path identical; the pipelined host-to-host call in the wide shape too; a sample against the oracle."""
    import bench

    jobs, _ = bench.build_workload(9000, 11, 0)
    out = {}
    for w in (4, 10):
        monkeypatch.setenv("PG2_LANE_W", str(w))
        with eng.batch(jobs) as b:
            b.run()
            st = eng.stats()
            out[w] = b.fetch()
        assert st["jobs_lanes"] == len(jobs) and st["jobs_lanes_wide"] == (len(jobs) if w == 10 else 0)
    monkeypatch.delenv("PG2_LANE_W")
    ra, sa = out[4]
    assert (ra["status"] == 0).all()
    piped = eng.align_prepared(eng.prepare(jobs, pinned=True, compact=True))  # >= 8192 jobs, few tasks: chunks in the wide shape
    assert eng.stats()["jobs_lanes_wide"] == len(jobs)
    for rb, sb in (out[10], piped):
        assert (rb["score"].view(np.uint64) == ra["score"].view(np.uint64)).all()
        assert (rb["status"] == ra["status"]).all() and (rb["n_steps"] == ra["n_steps"]).all()
        for k in range(0, len(jobs), 7):
            x = sb[rb["step_off"][k]: rb["step_off"][k] + rb["n_steps"][k]]
            y = sa[ra["step_off"][k]: ra["step_off"][k] + ra["n_steps"][k]]
            assert x.tobytes() == y.tobytes()
    heavy = sorted(range(len(jobs)), key=lambda k: -jobs[k].left.n_sites)[:3]  # the root-most targets
    for k in heavy + list(range(0, len(jobs), 3000)):
        job = enginecheck.expect_from_oracle(jobs[k])
        stp, _, _ = eng.expand(job, out[10][0][k], out[10][1])
        assert oracle_lib.steps_equal(stp, job.expected_path, job.expected_path_score) == []


def test_pipelined_call_equals_resident_batch_at_scale(eng, golden):
    """12 000 reads against 8 targets: pg2_align_batch takes its chunked, multi-context path by itself (>= 8192 jobs);
    every score, status and encoded path must equal what one resident batch (pg2_batch_create / run / fetch) gives, and
    a sample must equal the oracle."""
    model = golden["place_dna"][0].model
    rng = np.random.default_rng(4242)
    targets = [synth.leaf_graph(synth.random_dna(1500, rng)) for _ in range(8)]
    tseqs = [t.state[1:-1] for t in targets]
    reads, assign = [], []
    for k in range(12000):
        a = int(rng.integers(0, 8))
        st = int(rng.integers(0, 1350))
        r = tseqs[a][st:st + 150].copy()
        mut = rng.random(150) < 0.03
        r[mut] = rng.integers(0, 4, size=int(mut.sum()))
        reads.append(r)
        assign.append(a)
    jobs = synth.placement_jobs(targets, reads, assign, model)
    with eng.batch(jobs) as b:
        b.run()
        ra, sa = b.fetch()
    rb, sb = eng.align_prepared(eng.prepare(jobs, pinned=True, compact=True))
    rc, sc = eng.align(jobs)  # pageable buffers, explicit graphs
    for res, stp in ((rb, sb), (rc, sc)):
        assert (res["score"].view(np.uint64) == ra["score"].view(np.uint64)).all()
        assert (res["status"] == ra["status"]).all() and (res["n_steps"] == ra["n_steps"]).all()
        for k in range(0, len(jobs), 97):
            x = stp[res["step_off"][k]: res["step_off"][k] + res["n_steps"][k]]
            y = sa[ra["step_off"][k]: ra["step_off"][k] + ra["n_steps"][k]]
            assert x.tobytes() == y.tobytes()
    for k in range(0, len(jobs), 1500):
        job = enginecheck.expect_from_oracle(jobs[k])
        st, _, _ = eng.expand(job, rb[k], sb, compact=True)
        assert oracle_lib.steps_equal(st, job.expected_path, job.expected_path_score) == []


def test_band_kernel_vs_oracle(eng, eng_old):
    """Anchored leaf x leaf jobs through the band kernel and its segmented walk (see tests/test_emu_engine.py for the shapes),
    and the same jobs on the wavefront kernel's chain path."""
    import test_emu_engine

    jobs = test_emu_engine.band_jobs(401, reps=6)
    res = enginecheck.check_batch(eng, jobs)
    assert (res["kernel"][:-1] == 4).all() and res["kernel"][-1] == 0
    assert eng.stats()["jobs_band"] == len(jobs) - 1
    res = enginecheck.check_batch(eng_old, jobs)
    assert (res["kernel"] == 0).all()


@pytest.mark.parametrize("psring", [1, 0])
def test_pstrip_row_ring_and_register_rows(golden, psring):
    """Both step bodies of the pipelined-strip kernel (row ring in shared memory / row above in registers) on the reference's
    job streams at BASELINE sizes and on random general / banded graphs; see tests/test_emu_engine.py."""
    rng = np.random.default_rng(778)
    jobs = golden["c1_full"] + golden["c3_full"] + golden["c4_full"] + golden["c5_full"][2:] + golden["pileup_hp"] + golden["prog_dna"]
    jobs += [enginecheck.expect_from_oracle(randjobs.random_job(rng, kind)) for kind in ("general", "banded", "strip") for _ in range(100)]
    e = engine_with(PG2_NO_LANES=1, PG2_FORCE_PSRING=psring, PG2_NO_PSRING=1 - psring)
    try:
        res = enginecheck.check_batch(e, jobs)
        st = e.stats()
    finally:
        e.close()
    assert (res["kernel"] == 3).mean() > 0.9
    assert (st["jobs_pstrip_ring"] >= 0.9 * st["jobs_pstrip"]) if psring else (st["jobs_pstrip_ring"] == 0)
