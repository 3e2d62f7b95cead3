"""ctypes mirror of include/pagan2_b200.h (the C-ABI drop-in boundary).

Struct layouts and constants here must match the header field for field;
tests/test_abi.py checks the sizes against the compiled library.
"""
import ctypes as C
import os

import numpy as np

PG2_ABI_VERSION = 4

PG2_OK, PG2_ERR_INVALID, PG2_ERR_NO_DEVICE, PG2_ERR_CUDA, PG2_ERR_NOMEM, PG2_ERR_UNSUPPORTED, PG2_ERR_CAPACITY = range(7)
PG2_JOB_OK, PG2_JOB_NO_PATH, PG2_JOB_BAD_BAND, PG2_JOB_BAD_GRAPH, PG2_JOB_BROKEN_PATH = range(5)
PG2_X_MAT, PG2_Y_MAT, PG2_M_MAT = 0, 1, 2
PG2_FLAG_NO_TERMINAL_EDGES = 1
PG2_FLAG_REDUCED_TERMINAL_GAP_PENALTIES = 2
PG2_MAX_IN_DEGREE = 63

_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)


class Graph(C.Structure):
    _fields_ = [
        ("n_sites", C.c_int32),
        ("n_edges", C.c_int32),
        ("state", _i32p),
        ("bwd_off", _i32p),
        ("edge_start", _i32p),
        ("edge_logw", _f32p),
        ("edge_index", _i32p),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("fas", C.c_int32),
        ("log_score", _f32p),
        ("log_gap_open", C.c_float),
        ("log_gap_ext", C.c_float),
        ("log_gap_end_ext", C.c_float),
        ("log_gap_break_ext", C.c_float),
        ("log_non_gap", C.c_float),
    ]


class Job(C.Structure):
    _fields_ = [
        ("left", Graph),
        ("right", Graph),
        ("model", C.c_int32),
        ("flags", C.c_uint32),
        ("upper", _i32p),
        ("lower", _i32p),
    ]


class Result(C.Structure):
    _fields_ = [
        ("score", C.c_double),
        ("cells", C.c_int64),
        ("step_off", C.c_int64),
        ("n_steps", C.c_int32),
        ("status", C.c_int32),
        ("end_ptr", C.c_uint32),
        ("kernel", C.c_int32),
    ]


class Step(C.Structure):
    _fields_ = [
        ("score", C.c_double),
        ("matrix", C.c_int32),
        ("x_ind", C.c_int32),
        ("y_ind", C.c_int32),
        ("x_edge_ind", C.c_int32),
        ("y_edge_ind", C.c_int32),
        ("real_site", C.c_int32),
    ]


STEP_DTYPE = np.dtype(
    [
        ("score", "<f8"),
        ("matrix", "<i4"),
        ("x_ind", "<i4"),
        ("y_ind", "<i4"),
        ("x_edge_ind", "<i4"),
        ("y_edge_ind", "<i4"),
        ("real_site", "<i4"),
    ]
)
assert STEP_DTYPE.itemsize == C.sizeof(Step) == 32


class Stats(C.Structure):
    _fields_ = [
        ("fill_ms", C.c_double),
        ("traceback_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("cells", C.c_int64),
        ("traceback_bytes", C.c_int64),
        ("fill_launches", C.c_int32),
        ("traceback_launches", C.c_int32),
        ("jobs_wavefront", C.c_int32),
        ("jobs_strip", C.c_int32),
        ("run_ms", C.c_double),
        ("kernel_launches", C.c_int32),
        ("jobs_strip_groups", C.c_int32),
        ("jobs_lanes", C.c_int32),
        ("jobs_pstrip", C.c_int32),
        ("jobs_band", C.c_int32),
        ("jobs_pstrip_ring", C.c_int32),
        ("jobs_lanes_wide", C.c_int32),
    ]


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FlatGraph:
    """Owner of the numpy arrays behind one pg2_graph (keeps them alive)."""

    __slots__ = ("state", "off", "start", "logw", "eidx")

    def __init__(self, state, off, start, logw, eidx):
        self.state = np.ascontiguousarray(state, dtype=np.int32)
        self.off = np.ascontiguousarray(off, dtype=np.int32)
        self.start = np.ascontiguousarray(start, dtype=np.int32)
        self.logw = np.ascontiguousarray(logw, dtype=np.float32)
        self.eidx = np.ascontiguousarray(eidx, dtype=np.int32)
        assert self.off.shape[0] == self.state.shape[0] + 1
        assert self.start.shape[0] == self.logw.shape[0] == self.eidx.shape[0] == int(self.off[-1])

    @property
    def n_sites(self):
        return int(self.state.shape[0])

    def is_plain_chain(self):
        """Site s >= 1 entered by the one edge (s-1 -> s) with log weight +0.0: what the compact form describes."""
        n = self.n_sites
        if self.start.shape[0] != n - 1:
            return False
        return bool(self.off[0] == 0 and (self.off[1:] == np.arange(n, dtype=np.int32)).all()
                    and (self.start == np.arange(n - 1, dtype=np.int32)).all() and not self.logw.view(np.uint32).any())

    def as_struct(self, compact=False):
        """compact=True: a plain chain is passed in the compact form of pagan2_b200.h (states only; the edge indices stay
        for pg2_expand_path)."""
        g = Graph()
        g.n_sites = self.n_sites
        g.n_edges = int(self.start.shape[0])
        g.state = _ptr(self.state, C.c_int32)
        if compact and self.is_plain_chain():
            g.edge_index = _ptr(self.eidx, C.c_int32)
            return g
        g.bwd_off = _ptr(self.off, C.c_int32)
        g.edge_start = _ptr(self.start, C.c_int32)
        g.edge_logw = _ptr(self.logw, C.c_float)
        g.edge_index = _ptr(self.eidx, C.c_int32)
        return g

    @staticmethod
    def chain(states):
        """Plain leaf graph: start site, one site per state, stop site; in-degree 1, weight 1
        (reference: Sequence::create_default_sequence, sequence.cpp:152-303, default branch)."""
        n = len(states) + 2
        st = np.full(n, -1, np.int32)
        st[1:-1] = states
        off = np.zeros(n + 1, np.int32)
        off[1:] = np.arange(n, dtype=np.int32)  # site 0 has no backward edge, site s>=1 has edge s-1
        start = np.arange(0, n - 1, dtype=np.int32)
        logw = np.zeros(n - 1, np.float32)
        eidx = np.arange(1, n, dtype=np.int32)  # edge 0 is the reference's dummy first edge (-1 -> 0)
        return FlatGraph(st, off, start, logw, eidx)


class Model:
    """Owner of one pg2_model_desc."""

    def __init__(self, fas, table, scalars):
        self.fas = int(fas)
        self.table = np.ascontiguousarray(table, dtype=np.float32).reshape(-1)
        assert self.table.shape[0] == self.fas * self.fas
        self.scalars = np.ascontiguousarray(scalars, dtype=np.float32)
        assert self.scalars.shape[0] == 5

    def as_struct(self):
        m = ModelDesc()
        m.fas = self.fas
        m.log_score = _ptr(self.table, C.c_float)
        (m.log_gap_open, m.log_gap_ext, m.log_gap_end_ext, m.log_gap_break_ext, m.log_non_gap) = [
            float(x) for x in self.scalars
        ]
        return m


class FlatJob:
    """One alignment job in the flat layout plus (optionally) the reference's expected result."""

    def __init__(self, left, right, model, flags=PG2_FLAG_REDUCED_TERMINAL_GAP_PENALTIES, upper=None, lower=None):
        self.left, self.right, self.model, self.flags = left, right, model, int(flags)
        self.upper = None if upper is None else np.ascontiguousarray(upper, dtype=np.int32)
        self.lower = None if lower is None else np.ascontiguousarray(lower, dtype=np.int32)
        self.expected_score = None
        self.expected_path = None  # (n,6) int32: matrix,x_ind,y_ind,x_edge,y_edge,real
        self.expected_path_score = None
        self.meta = {}

    @property
    def cells(self):
        lx, ly = self.left.n_sites - 1, self.right.n_sites - 1
        if self.upper is None:
            return lx * ly
        lo = np.maximum(self.upper, 0)
        hi = np.minimum(self.lower, ly - 1)
        return int(np.maximum(hi - lo + 1, 0).sum())

    def as_struct(self, model_handle=0, compact=False):
        j = Job()
        j.left = self.left.as_struct(compact)
        j.right = self.right.as_struct(compact)
        j.model = model_handle
        j.flags = self.flags
        if self.upper is not None:
            j.upper = _ptr(self.upper, C.c_int32)
            j.lower = _ptr(self.lower, C.c_int32)
        return j
