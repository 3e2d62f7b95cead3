"""Shared checker: run a list of FlatJob through an Engine and compare with expectations/oracle."""
import numpy as np

import oracle_lib


def expect_from_oracle(job):
    status, score, steps, cells = oracle_lib.oracle_align(job)
    job.expected_status = status
    job.expected_score = score
    job.expected_path = np.stack([steps[n] for n in ("matrix", "x_ind", "y_ind", "x_edge_ind", "y_edge_ind", "real_site")],
                                 axis=1).astype(np.int32) if len(steps) else np.zeros((0, 6), np.int32)
    job.expected_path_score = steps["score"].copy()
    return job


def used_edges_from_path(job, path):
    """Edges the reference marks used, derived from the expected path (real steps carry the indices)."""
    real = path[path[:, 5] == 1]
    left = set(int(e) for e in real[:, 3] if e >= 0)
    right = set(int(e) for e in real[:, 4] if e >= 0)
    return left, right


def check_batch(eng, jobs, expect_kernel=None):
    """Returns the result array; asserts bit-exact score, identical path (every field, every per-step
    score bit) for each job."""
    res, steps = eng.align(jobs)
    assert len(res) == len(jobs)
    for k, job in enumerate(jobs):
        r = res[k]
        want_status = getattr(job, "expected_status", 0)
        assert r["status"] == want_status, "job %d: status %d, expected %d" % (k, r["status"], want_status)
        if want_status in (0, 1):
            assert r["cells"] == job.cells
        if want_status != 0:
            continue
        assert np.float64(r["score"]).view(np.uint64) == np.float64(job.expected_score).view(np.uint64), \
            "job %d: score %r vs %r" % (k, r["score"], job.expected_score)
        st, ul, ur = eng.expand(job, r, steps)
        diffs = oracle_lib.steps_equal(st, job.expected_path, job.expected_path_score)
        assert diffs == [], "job %d (kernel %d): %s" % (k, r["kernel"], diffs)
        wl, wr = used_edges_from_path(job, job.expected_path)
        assert wl <= set(ul.tolist()) and wr <= set(ur.tolist())
        if expect_kernel is not None:
            assert r["kernel"] == expect_kernel
    return res
