"""ctypes loader for the CPU checkers under oracle/ (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from pagan2_msa_b200 import abi

ORACLE_DIR = os.path.join(abi.REPO_ROOT, "oracle")
REF_SRC = "/root/reference/src"


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def build_ref():
    """Builds oracle/_ref when the reference sources are present (this container only)."""
    subprocess.check_call(["make", "-s", "-j8", "-C", ORACLE_DIR, "ref"])


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libviterbi_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        lib.pg2o_align.restype = C.c_int
        lib.pg2o_align.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.ModelDesc), C.POINTER(C.c_double),
                                   C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        lib.pg2o_align_marks.restype = C.c_int
        lib.pg2o_align_marks.argtypes = lib.pg2o_align.argtypes + [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int32)]
        _oracle = lib
    return _oracle


def oracle_align(job, marks=False):
    """Runs the C restatement on one FlatJob -> (status, score, steps[STEP_DTYPE], cells); with marks=True also the
    edge indices backtrack_new_path marks is_used(true), in marking order: (..., used_left, used_right)."""
    lib = oracle()
    cap = job.left.n_sites + job.right.n_sites + 2
    steps = np.zeros(cap, dtype=abi.STEP_DTYPE)
    js = job.as_struct()
    ms = job.model.as_struct()
    score = C.c_double()
    n = C.c_int32()
    cells = C.c_int64()
    ul, ur = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    nl, nr = C.c_int32(), C.c_int32()
    rc = lib.pg2o_align_marks(C.byref(js), C.byref(ms), C.byref(score), steps.ctypes.data, cap, C.byref(n), C.byref(cells),
                              ul.ctypes.data, C.byref(nl), ur.ctypes.data, C.byref(nr))
    if rc < 0:
        raise RuntimeError("oracle: capacity/allocation failure")
    if marks:
        return rc, score.value, steps[: n.value].copy(), cells.value, ul[: nl.value].copy(), ur[: nr.value].copy()
    return rc, score.value, steps[: n.value].copy(), cells.value


def ref_available():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libpagan2ref.so"))


def ref_binary():
    return os.path.join(ORACLE_DIR, "_ref", "pagan2_ref")


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libpagan2ref.so"))
        i32p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_float)
        lib.pagan2_ref_align_flat.restype = C.c_int
        lib.pagan2_ref_align_flat.argtypes = [C.c_int, f32p, f32p,
                                              C.c_int, i32p, i32p, i32p, f32p, i32p,
                                              C.c_int, i32p, i32p, i32p, f32p, i32p,
                                              i32p, i32p, C.c_int,
                                              C.POINTER(C.c_double), i32p, C.POINTER(C.c_double), C.c_int, i32p]
        lib.pagan2_ref_last_used.restype = C.c_int
        lib.pagan2_ref_last_used.argtypes = [C.c_int, i32p, C.c_int]
        if hasattr(lib, "pagan2_ref_prefix_anchors"):
            lib.pagan2_ref_prefix_anchors.restype = C.c_int
            lib.pagan2_ref_prefix_anchors.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, i32p, C.c_int]
            lib.pagan2_ref_anchor_band.restype = C.c_int
            lib.pagan2_ref_anchor_band.argtypes = [i32p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, i32p, i32p]
        _ref = lib
    return _ref


def ref_align_flat(job):
    """Runs the REFERENCE's Viterbi_alignment::align on one FlatJob -> (score, path(n,6), path_score)."""
    lib = ref_lib()
    cap = job.left.n_sites + job.right.n_sites + 2
    path = np.zeros(cap * 6, np.int32)
    ps = np.zeros(cap, np.float64)
    score = C.c_double()
    n = C.c_int32()
    p = abi._ptr
    L, R, m = job.left, job.right, job.model
    up = p(job.upper, C.c_int32) if job.upper is not None else None
    lo = p(job.lower, C.c_int32) if job.lower is not None else None
    rc = lib.pagan2_ref_align_flat(m.fas, p(m.table, C.c_float), p(m.scalars, C.c_float),
                                   L.n_sites, p(L.state, C.c_int32), p(L.off, C.c_int32), p(L.start, C.c_int32),
                                   p(L.logw, C.c_float), p(L.eidx, C.c_int32),
                                   R.n_sites, p(R.state, C.c_int32), p(R.off, C.c_int32), p(R.start, C.c_int32),
                                   p(R.logw, C.c_float), p(R.eidx, C.c_int32),
                                   up, lo, job.flags, C.byref(score), p(path, C.c_int32), p(ps, C.c_double), cap, C.byref(n))
    if rc != 0:
        raise RuntimeError("reference path longer than capacity")
    return score.value, path[: n.value * 6].reshape(-1, 6).copy(), ps[: n.value].copy()


def ref_prefix_anchors(seq1, seq2, min_length):
    """The REFERENCE's Find_anchors::find_long_substrings on two byte strings -> int32 (n, 3): start_1, start_2, length."""
    lib = ref_lib()
    cap = 4096
    while True:
        out = np.zeros(cap * 3, np.int32)
        n = lib.pagan2_ref_prefix_anchors(seq1, len(seq1), seq2, len(seq2), min_length, abi._ptr(out, C.c_int32), cap)
        if n <= cap:
            return out[: n * 3].reshape(-1, 3).copy()
        cap = n


def ref_anchor_band(hits, str1, str2, width):
    """The REFERENCE's Find_anchors::define_tunnel: hits (n, 3) + gapped strings -> (upper, lower)."""
    lib = ref_lib()
    hits = np.ascontiguousarray(hits, np.int32).reshape(-1)
    upper = np.zeros(len(str1) + 1, np.int32)
    lower = np.zeros(len(str1) + 1, np.int32)
    rc = lib.pagan2_ref_anchor_band(abi._ptr(hits, C.c_int32) if len(hits) else None, len(hits) // 3, str1, len(str1), str2, len(str2), width,
                                    abi._ptr(upper, C.c_int32), abi._ptr(lower, C.c_int32))
    if rc != 0:
        raise RuntimeError("reference define_tunnel returned vectors of unexpected length")
    return upper, lower


def ref_last_used():
    """Edge indices the last ref_align_flat call left marked is_used (fresh graphs: exactly the call's marks):
    (left, right), ascending."""
    lib = ref_lib()
    out = []
    for side in (0, 1):
        n = lib.pagan2_ref_last_used(side, None, 0)
        buf = np.zeros(max(n, 1), np.int32)
        lib.pagan2_ref_last_used(side, abi._ptr(buf, C.c_int32), n)
        out.append(buf[:n].copy())
    return out[0], out[1]


def steps_equal(steps, path, path_score, path_score_sha=None):
    """Bit-exact comparison of STEP_DTYPE steps with reference (n,6) path + scores (or, for long fixture paths, the
    SHA-256 of the scores' bytes); returns list of diffs."""
    diffs = []
    if len(steps) != len(path):
        return ["length %d vs %d" % (len(steps), len(path))]
    cols = ["matrix", "x_ind", "y_ind", "x_edge_ind", "y_edge_ind", "real_site"]
    for c, name in enumerate(cols):
        bad = np.nonzero(steps[name] != path[:, c])[0]
        if len(bad):
            diffs.append("%s differs at %d steps (first %d: %d vs %d)" % (name, len(bad), bad[0], steps[name][bad[0]], path[bad[0], c]))
    if path_score is not None:
        a = steps["score"].view(np.uint64)
        b = np.asarray(path_score, dtype=np.float64).view(np.uint64)
        bad = np.nonzero(a != b)[0]
        if len(bad):
            diffs.append("score bits differ at %d steps (first %d: %r vs %r)" % (len(bad), bad[0], steps["score"][bad[0]], path_score[bad[0]]))
    elif path_score_sha is not None:
        from pagan2_msa_b200 import jobio

        if jobio.sha_words(np.ascontiguousarray(steps["score"], dtype="<f8")).tolist() != np.asarray(path_score_sha).tolist():
            diffs.append("per-step scores: SHA-256 differs from the reference's")
    return diffs


def run_ref(args, cwd, dump_name="jobs.bin", stats_name="stats.json", timeout=3600):
    """Runs the reference binary (oracle/_ref/pagan2_ref) with the job-dump interposer on.
    Returns (list[FlatJob], stats dict)."""
    import json

    from pagan2_msa_b200 import jobio

    env = dict(os.environ)
    dump = os.path.join(cwd, dump_name)
    stats = os.path.join(cwd, stats_name)
    env["PAGAN2_ORACLE_DUMP"] = dump
    env["PAGAN2_ORACLE_STATS"] = stats
    subprocess.run([ref_binary()] + list(args), cwd=cwd, env=env, check=True, timeout=timeout,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    with open(stats) as f:
        st = json.load(f)
    return jobio.load_jobs(dump), st
