"""world_size-2 test of the multi-GPU path (pagan2_msa_b200/shard.py) on CPU: two processes over gloo, each
running its index-range shard through the CPU test build of the engine (tests/_emu), results gathered to
rank 0 and compared bit for bit with the single-rank run and with the reference's golden dump."""
import os
import subprocess
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (_ROOT, os.path.join(_ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import enginecheck  # noqa: E402
import oracle_lib  # noqa: E402
from pagan2_msa_b200 import abi, engine, jobio, shard  # noqa: E402

EMU_DIR = os.path.join(abi.REPO_ROOT, "tests", "emu")
EMU_LIB = os.path.join(abi.REPO_ROOT, "tests", "_emu", "libpg2_emu.so")


def test_partition_is_a_balanced_permutation():
    rng = np.random.default_rng(5)
    cells = rng.integers(1, 10**6, size=1001)
    for world in (1, 2, 3, 8):
        parts = shard.partition(cells, world)
        assert len(parts) == world
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(1001))
        loads = [int(cells[p].sum()) for p in parts]
        assert max(loads) - min(loads) <= cells.max()
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1
    assert [len(p) for p in shard.partition([], 4)] == [0, 0, 0, 0]
    assert [len(p) for p in shard.partition([7], 4)] == [1, 0, 0, 0]


def test_partition_keeps_groups_together():
    """groups=: the reads of one target stay on one rank except at the rank borders, units of 32 never mix two targets, the
    cell counts balance to within one unit, and the result does not depend on the key values."""
    rng = np.random.default_rng(6)
    n, ng = 20000, 127
    g = rng.integers(0, ng, size=n)
    lens = rng.integers(1400, 3000, size=ng)
    cells = lens[g] * rng.integers(140, 152, size=n)
    for world in (1, 2, 4, 8):
        parts = shard.partition(cells, world, groups=g.tolist())
        assert len(parts) == world
        assert sorted(np.concatenate(parts).tolist()) == list(range(n))
        loads = [int(cells[p].sum()) for p in parts]
        assert max(loads) - min(loads) <= 2 * 32 * int(cells.max())
        owners = {}
        for r, p in enumerate(parts):
            for k in set(g[p].tolist()):
                owners.setdefault(k, []).append(r)
        assert sum(len(v) > 1 for v in owners.values()) <= world - 1  # only a group at a border is cut
        tasks = sum(int(((np.bincount(g[p], minlength=ng) + 31) // 32).sum()) for p in parts)
        assert tasks <= int(((np.bincount(g, minlength=ng) + 31) // 32).sum()) + (world - 1)
        same = shard.partition(cells, world, groups=["t%d" % (ng - k) for k in g.tolist()])
        assert all((a == b).all() for a, b in zip(parts, same))
    assert [len(p) for p in shard.partition([], 3, groups=[])] == [0, 0, 0]
    assert sorted(np.concatenate(shard.partition([5, 3, 9], 2, groups=["a", "b", "a"])).tolist()) == [0, 1, 2]


def test_placement_shards_balance_estimated_time():
    """placement_costs weighs the virtual rows of multi-edge sites; placement_shards cuts whole targets to equal cost."""
    jobs = jobio.load_jobs(os.path.join(abi.REPO_ROOT, "tests", "golden", "place_dna.pjob.gz"))
    costs = shard.placement_costs(jobs)
    assert costs.shape[0] == len(jobs) and (costs > 0).all()
    for j, c in zip(jobs, costs):
        plain_rows = j.left.n_sites  # a general site costs at least its one virtual row
        assert c >= (plain_rows - 1) * max(j.right.n_sites - 1, 1) * 0.99
        if j.left.is_plain_chain():
            assert c == j.left.n_sites * max(j.right.n_sites - 1, 1)
    for world in (1, 2, 3):
        parts = shard.placement_shards(jobs, world)
        assert sorted(np.concatenate(parts).tolist()) == list(range(len(jobs)))


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    jobs = jobio.load_jobs(os.path.join(abi.REPO_ROOT, "tests", "golden", "place_dna.pjob.gz"))[:24]
    jobs += jobio.load_jobs(os.path.join(abi.REPO_ROOT, "tests", "golden", "anchored.pjob.gz"))[:3]
    with engine.Engine(0, EMU_LIB) as eng:
        got = shard.align_sharded(eng, jobs, dist, rank, world, torch.device("cpu"))
        if rank == 0:
            records, step_off, steps = got
            res = engine.results_from_records(records, step_off, jobs)
            ref_res, ref_steps = eng.align(jobs)
            assert (res["score"].view(np.uint64) == ref_res["score"].view(np.uint64)).all()
            assert (res["status"] == ref_res["status"]).all() and (res["n_steps"] == ref_res["n_steps"]).all()
            for k, job in enumerate(jobs):
                a, ula, ura = eng.expand(job, res[k], steps)
                b, ulb, urb = eng.expand(job, ref_res[k], ref_steps)
                assert a.tobytes() == b.tobytes() and ula.tolist() == ulb.tolist() and ura.tolist() == urb.tolist()
                assert oracle_lib.steps_equal(a, job.expected_path, job.expected_path_score) == []
            open(out_path, "w").write("ok %d" % len(jobs))
        else:
            assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    out = str(tmp_path / "rank0.txt")
    port = 29500 + os.getpid() % 2000
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), str(r), "2", str(port), out]) for r in range(2)]
    codes = [p.wait(timeout=600) for p in procs]
    assert codes == [0, 0]
    assert open(out).read().startswith("ok 27")


if __name__ == "__main__":
    _worker(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
