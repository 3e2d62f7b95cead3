"""Shared checker: run a list of FlatJob through an Engine and compare with expectations/oracle."""
import numpy as np

import oracle_lib


def expect_from_oracle(job):
    status, score, steps, cells, ul, ur = oracle_lib.oracle_align(job, marks=True)
    job.oracle_marks = (ul, ur)
    job.expected_status = status
    job.expected_score = score
    job.expected_path = np.stack([steps[n] for n in ("matrix", "x_ind", "y_ind", "x_edge_ind", "y_edge_ind", "real_site")],
                                 axis=1).astype(np.int32) if len(steps) else np.zeros((0, 6), np.int32)
    job.expected_path_score = steps["score"].copy()
    return job


def check_marks(job, ul, ur, k=0):
    """The is_used(true) marks pg2_expand_path replays (viterbi_alignment.cpp:1054-1155): equal, in order, to the oracle's
    restatement of backtrack_new_path; and as a set equal to what the REFERENCE itself left marked (fixtures dumped by the
    interposer carry the flags before and after the reference's call)."""
    marks = getattr(job, "oracle_marks", None)
    if marks is None:
        marks = oracle_lib.oracle_align(job, marks=True)[4:6]
    assert ul.tolist() == marks[0].tolist(), "job %d: left edge marks differ from the oracle's (order included)" % k
    assert ur.tolist() == marks[1].tolist(), "job %d: right edge marks differ from the oracle's (order included)" % k
    used = getattr(job, "expected_used", None)
    if used is not None:
        for side, mine in (("l", ul), ("r", ur)):
            before, after = set(used[side][0].tolist()), set(used[side][1].tolist())
            assert before | set(mine.tolist()) == after, "job %d: %s edge marks differ from the reference's" % (k, side)


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
        diffs = oracle_lib.steps_equal(st, job.expected_path, job.expected_path_score, getattr(job, 'expected_path_score_sha', None))
        assert diffs == [], "job %d (kernel %d): %s" % (k, r["kernel"], diffs)
        check_marks(job, ul, ur, k)
        if expect_kernel is not None:
            assert r["kernel"] == expect_kernel
    return res
