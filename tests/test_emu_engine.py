"""CPU tests of the engine's HOST logic (packing, graph sharing, band geometry, launch grouping, result
unpacking) and of the per-thread kernel bodies, through tests/_emu/libpg2_emu.so: the product's CUDA
sources compiled with g++ against tests/emu/pg2_emu_runtime.h, kernels run as serial loops.  This is a
development aid for a container without a GPU -- the shipped library is built by nvcc and has no CPU
path; the GPU parity tests proper are in test_gpu_parity.py."""
import os
import subprocess

import numpy as np
import pytest

import enginecheck
import randjobs
from pagan2_msa_b200 import abi, engine

EMU_DIR = os.path.join(abi.REPO_ROOT, "tests", "emu")
EMU_LIB = os.path.join(abi.REPO_ROOT, "tests", "_emu", "libpg2_emu.so")


@pytest.fixture(scope="module")
def emu_lib():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    return EMU_LIB


def make_engine(lib, force_wavefront, no_lanes=False, no_pstrip=False, pstrip_k=None, no_band=False, psring=None):
    """no_pstrip: keep the pipelined-strip kernel out, so that the warp-per-alignment strip kernel and the general
    wavefront kernel (its fallback) stay covered; pstrip_k=4 forces the wider strips."""
    os.environ["PG2_FORCE_WAVEFRONT"] = "1" if force_wavefront else "0"
    os.environ["PG2_NO_LANES"] = "1" if no_lanes else "0"
    os.environ["PG2_NO_PSTRIP"] = "1" if no_pstrip else "0"
    os.environ["PG2_NO_BAND"] = "1" if no_band else "0"
    os.environ["PG2_FORCE_PSRING"] = "1" if psring is True else "0"  # the row-ring step for every eligible pipelined-strip job
    os.environ["PG2_NO_PSRING"] = "1" if psring is False else "0"    # ... for none
    if pstrip_k:
        os.environ["PG2_PSTRIP_K"] = str(pstrip_k)
        os.environ["PG2_PSTRIP_BANDED_CHAINS"] = "1"  # (by default banded chain x chain jobs stay on the wavefront kernel)
    try:
        return engine.Engine(0, lib)
    finally:
        for name in ("PG2_FORCE_WAVEFRONT", "PG2_NO_LANES", "PG2_NO_PSTRIP", "PG2_PSTRIP_K", "PG2_PSTRIP_BANDED_CHAINS", "PG2_NO_BAND", "PG2_FORCE_PSRING", "PG2_NO_PSRING"):
            os.environ.pop(name, None)


@pytest.mark.parametrize("force_wavefront", [True, False])
@pytest.mark.parametrize("name", ["prog_dna", "place_dna", "pileup_hp", "codon", "anchored"])
def test_golden(emu_lib, golden, name, force_wavefront):
    with make_engine(emu_lib, force_wavefront) as eng:
        res = enginecheck.check_batch(eng, golden[name])
        if force_wavefront:
            assert (res["kernel"] == 0).all()


def test_strip_kernel_is_chosen_for_placement(emu_lib, golden):
    with make_engine(emu_lib, False, no_lanes=True, no_pstrip=True) as eng:
        res = enginecheck.check_batch(eng, golden["place_dna"])
        assert (res["kernel"] == 1).all()
    with make_engine(emu_lib, False, no_pstrip=True) as eng:
        res = enginecheck.check_batch(eng, golden["place_dna"])
        assert np.isin(res["kernel"], (1, 2)).all()
    with make_engine(emu_lib, False, no_lanes=True) as eng:  # a batch this small goes to the CTA-per-alignment kernel
        res = enginecheck.check_batch(eng, golden["place_dna"])
        assert (res["kernel"] == 3).all()


@pytest.mark.parametrize("name", ["prog_dna", "place_dna", "pileup_hp", "codon", "anchored"])
@pytest.mark.parametrize("k", [2, 4])
def test_golden_pstrip(emu_lib, golden, name, k):
    """Every reference job stream through the pipelined-strip kernel (general graphs on both sides, bands), both strip
    widths; and with the older kernels only."""
    with make_engine(emu_lib, False, no_lanes=True, pstrip_k=k) as eng:
        res = enginecheck.check_batch(eng, golden[name])
        assert (res["kernel"] == 3).all()
    with make_engine(emu_lib, False, no_pstrip=True) as eng:
        res = enginecheck.check_batch(eng, golden[name])
        assert (res["kernel"] != 3).all()


@pytest.mark.parametrize("seed,plain_left,n_jobs", [(61, False, 70), (62, True, 40), (63, False, 33), (64, False, 16), (65, True, 100)])
@pytest.mark.parametrize("lane_w", [4, 10])  # three CTAs of 4 warps per SM / one CTA of 10: the throughput and the latency shape
def test_lane_kernel_shared_target_vs_oracle(emu_lib, monkeypatch, lane_w, seed, plain_left, n_jobs):
    """Jobs that share the left graph are grouped 32 per warp, one alignment per lane (pg2_lanes.cu): every
    variant (plain / general rows, weights on the read edges), ragged read lengths, thin remainders."""
    monkeypatch.setenv("PG2_LANE_W", str(lane_w))
    rng = np.random.default_rng(seed)
    jobs = []
    for _ in range(3):
        jobs += randjobs.random_shared_target_jobs(rng, n_jobs, plain_left=plain_left)
    jobs += [randjobs.random_job(rng, "strip") for _ in range(5)]  # singletons stay on the strip kernel
    jobs = [enginecheck.expect_from_oracle(j) for j in jobs]
    with make_engine(emu_lib, False, no_pstrip=True) as eng:
        res = enginecheck.check_batch(eng, jobs)
        st = eng.stats()
    assert (res["kernel"][:3 * n_jobs] == 2).all()
    assert (res["kernel"][-5:] == 1).all()
    assert st["jobs_lanes"] == int((res["kernel"] == 2).sum())
    assert st["jobs_lanes_wide"] == (st["jobs_lanes"] if lane_w == 10 else 0)


@pytest.mark.parametrize("seed", [71, 72, 73, 74])
@pytest.mark.parametrize("lane_w", [4, 10])
def test_lane_kernel_schedule_shapes(emu_lib, monkeypatch, lane_w, seed):
    """Pipeline schedule edge cases: row programs shorter than the pipeline depth, one strip only, many
    rounds (reads of several hundred columns), reads of one site."""
    monkeypatch.setenv("PG2_LANE_W", str(lane_w))
    rng = np.random.default_rng(seed)
    jobs = []
    for nl, nr_max in ((1, 4), (2, 9), (5, 40), (30, 9), (33, 300), (150, 120), (9, 70)):
        jobs += randjobs.random_shared_target_jobs(rng, 20, nl=nl, nr_max=nr_max)
    jobs = [enginecheck.expect_from_oracle(j) for j in jobs]
    with make_engine(emu_lib, False) as eng:
        res = enginecheck.check_batch(eng, jobs)
    assert (res["kernel"] == 2).all()


@pytest.mark.parametrize("lane_w", [4, 10])
def test_lane_kernel_long_reads_and_bad_job(emu_lib, monkeypatch, lane_w):
    """Reads longer than one strip x many strips, and a rejected job inside a task (its lane idles)."""
    monkeypatch.setenv("PG2_LANE_W", str(lane_w))
    rng = np.random.default_rng(66)
    jobs = randjobs.random_shared_target_jobs(rng, 40, nl=60, nr_max=200)
    jobs = [enginecheck.expect_from_oracle(j) for j in jobs]
    jobs[7].right.state[1] = 99
    jobs[7].expected_status = abi.PG2_JOB_BAD_GRAPH
    with make_engine(emu_lib, False) as eng:
        res = enginecheck.check_batch(eng, jobs)
    assert (res["kernel"] == 2).all()


@pytest.mark.parametrize("kind,seed", [("general", 21), ("banded", 22), ("strip", 23), ("banded_chain", 24)])
def test_random_jobs_vs_oracle(emu_lib, kind, seed):
    rng = np.random.default_rng(seed)
    jobs = [enginecheck.expect_from_oracle(randjobs.random_job(rng, kind)) for _ in range(60)]
    with make_engine(emu_lib, False) as eng:
        enginecheck.check_batch(eng, jobs)


def test_shared_graphs_uploaded_once(emu_lib, golden):
    jobs = golden["place_dna"]
    # 60 jobs, but far fewer distinct target graphs
    with make_engine(emu_lib, False) as eng:
        eng.align(jobs + jobs)
        once = eng.stats()["h2d_bytes"]
        eng.align(jobs)
        assert eng.stats()["h2d_bytes"] < once  # doubling the jobs re-used every graph: only job records grew
        assert once < 2 * eng.stats()["h2d_bytes"]


def test_small_scratch_budget_splits_groups(emu_lib, golden):
    os.environ["PG2_SCRATCH_MB"] = "1"
    try:
        with make_engine(emu_lib, True) as eng:
            enginecheck.check_batch(eng, golden["prog_dna"])
            assert eng.stats()["fill_launches"] > 1
    finally:
        os.environ.pop("PG2_SCRATCH_MB", None)


def test_bad_inputs_are_reported_not_crashed(emu_lib):
    rng = np.random.default_rng(31)
    good = enginecheck.expect_from_oracle(randjobs.random_job(rng, "general"))
    bad_graph = randjobs.random_job(rng, "general")
    bad_graph.left.start[-1] = bad_graph.left.n_sites + 5  # edge from a later site
    bad_graph.expected_status = abi.PG2_JOB_BAD_GRAPH
    bad_state = randjobs.random_job(rng, "general")
    bad_state.right.state[1] = 99
    bad_state.expected_status = abi.PG2_JOB_BAD_GRAPH
    bad_band = randjobs.random_job(rng, "banded")
    bad_band.upper = bad_band.upper.copy()
    bad_band.upper[len(bad_band.upper) // 2] = bad_band.upper[-1] + 5
    bad_band.expected_status = abi.PG2_JOB_BAD_BAND
    no_path = randjobs.random_job(rng, "general")
    while no_path.right.n_sites < 6:
        no_path = randjobs.random_job(rng, "general")
    lx = no_path.left.n_sites - 1
    no_path.upper = np.zeros(lx, np.int32)
    no_path.lower = np.zeros(lx, np.int32)
    no_path.expected_status = abi.PG2_JOB_NO_PATH
    with make_engine(emu_lib, False) as eng:
        enginecheck.check_batch(eng, [good, bad_graph, bad_state, bad_band, no_path, good])


def test_empty_batch_and_tiny_graphs(emu_lib):
    rng = np.random.default_rng(41)
    with make_engine(emu_lib, False) as eng:
        res, steps = eng.align([])
        assert len(res) == 0
        model = randjobs.random_model(rng, 15)
        empty = abi.FlatGraph.chain(np.zeros(0, np.int32))  # start + stop only
        one = abi.FlatGraph.chain(np.array([2], np.int32))
        jobs = [abi.FlatJob(empty, empty, model, 2), abi.FlatJob(one, empty, model, 2), abi.FlatJob(empty, one, model, 3),
                abi.FlatJob(one, one, model, 0)]
        jobs = [enginecheck.expect_from_oracle(j) for j in jobs]
        enginecheck.check_batch(eng, jobs)


def test_invalid_arguments(emu_lib):
    with make_engine(emu_lib, False) as eng:
        rng = np.random.default_rng(51)
        job = randjobs.random_job(rng, "general")
        js = job.as_struct(12345)  # unknown model handle
        import ctypes as C
        arr = (abi.Job * 1)(js)
        res = (abi.Result * 1)()
        steps = np.zeros(1000, np.uint16)
        rc = eng.lib.pg2_align_batch(eng.ctx, 1, arr, res, steps.ctypes.data, 1000)
        assert rc == abi.PG2_ERR_INVALID
        assert b"model" in eng.lib.pg2_last_error()


@pytest.mark.parametrize("chunks", [2, 3, 5])
def test_pipelined_align_batch_matches_single_shot(emu_lib, chunks):
    """pg2_align_batch cuts large batches into chunks that alternate between two contexts (host packing of chunk
    k+1 overlaps the device work of chunk k).  Same results, same packed paths per job, any chunk count; jobs that
    share a left graph stay together."""
    rng = np.random.default_rng(80 + chunks)
    jobs = []
    for _ in range(4):
        jobs += randjobs.random_shared_target_jobs(rng, 37)
    jobs += [randjobs.random_job(rng, kind) for kind in ("strip", "general", "banded") for _ in range(6)]
    rng.shuffle(jobs)
    jobs = [enginecheck.expect_from_oracle(j) for j in jobs]
    os.environ["PG2_PIPELINE_MIN_JOBS"] = "8"
    os.environ["PG2_PIPELINE_CHUNKS"] = str(chunks)
    try:
        with make_engine(emu_lib, False) as eng:
            res = enginecheck.check_batch(eng, jobs)
            assert (res["kernel"] == 2).sum() >= 4 * 32
            assert sum(eng.stats()[k] for k in ("jobs_lanes", "jobs_strip", "jobs_wavefront", "jobs_pstrip", "jobs_band")) == len(jobs)
            # twice on the same engine: the sibling context and its buffers are reused
            enginecheck.check_batch(eng, jobs[::-1])
    finally:
        os.environ.pop("PG2_PIPELINE_MIN_JOBS", None)
        os.environ.pop("PG2_PIPELINE_CHUNKS", None)


def test_pipelined_align_batch_dominant_last_job(emu_lib):
    """A batch of short alignments plus one that holds most of the cells, last in left-graph order: the chunk boundary search
    lands behind the last job (it used to read one past the permutation there)."""
    rng = np.random.default_rng(95)
    jobs = randjobs.random_shared_target_jobs(rng, 40, nl=30, nr_max=40)
    big = randjobs.random_job(rng, "general")
    while big.cells < 4 * sum(j.cells for j in jobs):
        big = abi.FlatJob(randjobs.random_graph(rng, 400, 4, p_extra=0.05), randjobs.random_graph(rng, 400, 4, p_extra=0.05), jobs[0].model, 2)
    jobs = [enginecheck.expect_from_oracle(j) for j in jobs + [big]]
    os.environ["PG2_PIPELINE_MIN_JOBS"] = "8"
    os.environ["PG2_PIPELINE_CHUNKS"] = "4"
    try:
        with make_engine(emu_lib, False) as eng:
            res = enginecheck.check_batch(eng, jobs)
            assert (res["status"] == 0).all()
    finally:
        os.environ.pop("PG2_PIPELINE_MIN_JOBS", None)
        os.environ.pop("PG2_PIPELINE_CHUNKS", None)


def test_compact_chain_form(emu_lib, golden):
    """Plain chains passed in the compact form of pagan2_b200.h (states only, CSR arrays NULL): same results, same packed
    paths, same expanded paths as the explicit form; fewer bytes staged; a compact graph that is no chain is refused."""
    rng = np.random.default_rng(91)
    jobs = list(golden["place_dna"][:24]) + randjobs.random_shared_target_jobs(rng, 40, weights=False)
    jobs += [randjobs.random_job(rng, kind) for kind in ("banded_chain", "strip") for _ in range(8)]
    jobs += randjobs.random_shared_target_jobs(rng, 20, plain_left=True, weights=False)
    assert sum(j.right.is_plain_chain() for j in jobs) > 80 and any(j.left.is_plain_chain() for j in jobs)
    assert any(not j.right.is_plain_chain() for j in jobs)  # weighted chains keep the explicit form
    with make_engine(emu_lib, False) as eng:
        ra, sa = eng.align(jobs)
        ra, sa = ra.copy(), sa.copy()
        bytes_explicit = eng.stats()["h2d_bytes"]
        rb, sb = eng.align_prepared(eng.prepare(jobs, compact=True))
        assert eng.stats()["h2d_bytes"] <= bytes_explicit
        assert ra.tobytes() == rb.tobytes()
        for k, job in enumerate(jobs):
            if ra["status"][k] != 0:
                continue
            pa, la, ua = eng.expand(job, ra[k], sa)
            pb, lb, ub = eng.expand(job, rb[k], sb, compact=True)
            assert pa.tobytes() == pb.tobytes() and la.tobytes() == lb.tobytes() and ua.tobytes() == ub.tobytes()
        # n_edges must be n_sites - 1 in the compact form
        bad = jobs[0].as_struct(eng.model_handle(jobs[0].model), compact=True)
        assert not bad.right.bwd_off
        bad.right.n_edges += 1
        arr = (abi.Job * 1)(bad)
        res = (abi.Result * 1)()
        stp = np.zeros(4096, np.uint16)
        assert eng.lib.pg2_align_batch(eng.ctx, 1, arr, res, stp.ctypes.data, 4096) == abi.PG2_ERR_INVALID


def test_path_runs_longer_than_one_repeat_word(emu_lib):
    """A 40 000-site target against a 5-nt read: the walk holds one gap run longer than the 32 767 a repeat word can
    carry; the encoded path is a few words and expands to the oracle's path."""
    rng = np.random.default_rng(404)
    model = randjobs.random_model(rng, 4)
    left = abi.FlatGraph.chain(rng.integers(0, 4, size=40000).astype(np.int32))
    right = abi.FlatGraph.chain(rng.integers(0, 4, size=5).astype(np.int32))
    job = enginecheck.expect_from_oracle(abi.FlatJob(left, right, model, 2))
    with make_engine(emu_lib, False) as eng:
        res = enginecheck.check_batch(eng, [job])
        assert res["n_steps"][0] < 64 and len(job.expected_path) > 39000


@pytest.mark.parametrize("kind,seed,k", [("general", 81, 2), ("general", 82, 4), ("banded", 83, 2), ("banded", 84, 4),
                                         ("strip", 85, 2), ("banded_chain", 86, 2)])
def test_pstrip_random_jobs_vs_oracle(emu_lib, kind, seed, k):
    """Pipelined-strip kernel on seeded random jobs: multi-edge graphs with long spans and weights on both sides, tie-prone
    parameter sets, random monotone bands, every flag combination."""
    rng = np.random.default_rng(seed)
    jobs = [enginecheck.expect_from_oracle(randjobs.random_job(rng, kind)) for _ in range(60)]
    with make_engine(emu_lib, False, no_lanes=True, pstrip_k=k) as eng:
        res = enginecheck.check_batch(eng, jobs)
    assert (res["kernel"] == 3).mean() > 0.9


@pytest.mark.parametrize("seed,banded", [(91, False), (92, True), (93, False), (94, True)])
def test_pstrip_many_blocks(emu_lib, seed, banded):
    """Jobs wide enough for several column blocks per alignment (the blocks of the CTA pipeline): general x general graphs
    of a few hundred sites, long-span edges that cross lanes, parked rows read across a block's lanes, bands that leave
    some blocks without rows."""
    rng = np.random.default_rng(seed)
    jobs = []
    for _ in range(6):
        model = randjobs.random_model(rng, 15, ties=rng.random() < 0.4)
        nl, nr = int(rng.integers(150, 420)), int(rng.integers(150, 420))
        left = randjobs.random_graph(rng, nl, 15, p_extra=0.1, max_span=int(rng.integers(3, 25)))
        right = randjobs.random_graph(rng, nr, 15, p_extra=float(rng.choice([0.03, 0.12])), max_span=int(rng.integers(3, 20)))
        job = abi.FlatJob(left, right, model, int(rng.integers(0, 4)))
        if banded:
            job.upper, job.lower = randjobs.random_band(rng, left.n_sites - 1, right.n_sites - 1, min_w=3, max_w=40)
        jobs.append(enginecheck.expect_from_oracle(job))
    with make_engine(emu_lib, False) as eng:
        res = enginecheck.check_batch(eng, jobs)
        st = eng.stats()
    # (a right graph without a cut point within a block's width stays on the general wavefront kernel)
    assert (res["kernel"] == 3).sum() >= len(jobs) - 1 and st["jobs_pstrip"] == int((res["kernel"] == 3).sum())


@pytest.mark.parametrize("name", ["c1_full", "c3_full", "c5_full"])
def test_reference_streams_at_baseline_size(emu_lib, golden, name):
    """The engine's host logic and the pipelined-strip kernel body on job streams the reference ran at BASELINE sizes
    (several column blocks per alignment, real anchor bands, 200 kb graphs)."""
    jobs = golden[name] if name != "c3_full" else golden[name][::3]
    with make_engine(emu_lib, False) as eng:
        res = enginecheck.check_batch(eng, jobs)
    assert (res["kernel"] == 3).sum() >= len(jobs) - 2


BAND_SHAPES = [(1, {}), (2, {}), (5, {}), (40, {}), (300, {}), (700, dict(bulge=40)), (1500, dict(bulge=120, width=(8, 30))),
               (900, dict(fas=211)), (600, dict(disconnect=True)), (1200, dict(bulge=300, width=(10, 40)))]


def band_jobs(seed, reps=2):
    rng = np.random.default_rng(seed)
    jobs = [enginecheck.expect_from_oracle(randjobs.random_anchor_band_job(rng, n, **kw)) for n, kw in BAND_SHAPES for _ in range(reps)]
    # wider than the band kernel's row ring: the wavefront kernel keeps the job
    wide = randjobs.random_anchor_band_job(rng, 600)
    wide.upper = np.zeros_like(wide.upper)
    wide.lower = np.full_like(wide.lower, wide.right.n_sites + 2)
    jobs.append(enginecheck.expect_from_oracle(wide))
    return jobs


def test_band_kernel_vs_oracle(emu_lib):
    """Anchored leaf x leaf jobs through the band kernel (one warp per job, rings indexed by row, M one diagonal ahead) and
    its segmented walk: one-site graphs, several 32-step chunks and walk segments, diagonals longer than a warp (several
    passes per step), the table outside shared memory (fas 211), a band cut in two (no path), every flag combination."""
    jobs = band_jobs(301)
    with make_engine(emu_lib, False) as eng:
        res = enginecheck.check_batch(eng, jobs)
        assert (res["kernel"][:-1] == 4).all() and res["kernel"][-1] == 0
        assert eng.stats()["jobs_band"] == len(jobs) - 1
    with make_engine(emu_lib, False, no_band=True) as eng:  # the same jobs on the wavefront kernel's chain path
        res = enginecheck.check_batch(eng, jobs)
        assert (res["kernel"] == 0).all()


@pytest.mark.parametrize("psring", [True, False])
def test_pstrip_row_ring_and_register_rows(emu_lib, golden, psring):
    """The two step bodies of the pipelined-strip kernel -- every source row from the shared-memory row ring (ps_step_ring) and
    the row above in registers with parked rows in global memory (ps_step) -- on the reference's own job streams (ancestors,
    pileup root with 63-site edges: the 128-row ring, 1892-state codons) and on random general / banded graphs."""
    rng = np.random.default_rng(777)
    jobs = golden["c1_full"][8:] + golden["c3_full"][-6:] + golden["c4_full"][-2:] + golden["pileup_hp"] + golden["prog_dna"]
    jobs += [enginecheck.expect_from_oracle(randjobs.random_job(rng, kind)) for kind in ("general", "banded", "strip") for _ in range(25)]
    with make_engine(emu_lib, False, no_lanes=True, psring=psring) as eng:
        res = enginecheck.check_batch(eng, jobs)
        st = eng.stats()
    assert (res["kernel"] == 3).mean() > 0.9
    if psring:
        assert st["jobs_pstrip_ring"] >= 0.9 * st["jobs_pstrip"]
    else:
        assert st["jobs_pstrip_ring"] == 0
