"""Host-side Python binding of the C-ABI (include/pagan2_b200.h) used by tests and bench.py.

`Engine` wraps one pg2_ctx.  The product library is pagan2_msa_b200/libpagan2_b200.so (nvcc, sm_100a);
if it is missing or no B200 is visible every call raises -- there is no CPU path behind this class.
"""
import ctypes as C
import os

import numpy as np

from . import abi

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpagan2_b200.so")


class Pg2Error(RuntimeError):
    def __init__(self, code, text):
        super().__init__("pg2 error %d: %s" % (code, text))
        self.code = code


def _bind(lib):
    vp = C.c_void_p
    lib.pg2_abi_version.restype = C.c_int
    lib.pg2_last_error.restype = C.c_char_p
    lib.pg2_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.pg2_ctx_destroy.argtypes = [vp]
    lib.pg2_ctx_destroy.restype = None
    lib.pg2_model_upload.argtypes = [vp, C.POINTER(abi.ModelDesc), C.POINTER(C.c_int32)]
    lib.pg2_model_release.argtypes = [vp, C.c_int32]
    lib.pg2_align_batch.argtypes = [vp, C.c_int32, C.POINTER(abi.Job), C.POINTER(abi.Result), vp, C.c_int64]
    lib.pg2_batch_create.argtypes = [vp, C.c_int32, C.POINTER(abi.Job), C.POINTER(vp)]
    lib.pg2_batch_run.argtypes = [vp, vp]
    lib.pg2_batch_fetch.argtypes = [vp, vp, C.POINTER(abi.Result), vp, C.c_int64]
    lib.pg2_batch_step_capacity.argtypes = [vp]
    lib.pg2_batch_step_capacity.restype = C.c_int64
    lib.pg2_batch_destroy.argtypes = [vp, vp]
    lib.pg2_batch_destroy.restype = None
    lib.pg2_expand_path.argtypes = [C.POINTER(abi.Job), C.POINTER(abi.ModelDesc), C.POINTER(abi.Result), vp, vp,
                                    C.POINTER(C.c_int32), vp, C.POINTER(C.c_int32), vp, C.POINTER(C.c_int32)]
    lib.pg2_get_stats.argtypes = [vp, C.POINTER(abi.Stats)]
    lib.pg2_batch_device_buffers.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int64)]
    lib.pg2_stream_synchronize.argtypes = [vp]
    lib.pg2_measure_fp64_issue.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.pg2_measure_dispatch_mix.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.pg2_find_prefix_anchors.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, vp, C.c_int32, C.POINTER(C.c_int32)]
    lib.pg2_anchor_band.argtypes = [vp, C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, vp, vp]
    return lib


_libs = {}


def load_library(path=None):
    path = path or os.environ.get("PG2_LIB") or LIB_PATH  # PG2_LIB: an alternative build of the same CUDA library (tuning)
    if path not in _libs:
        if not os.path.exists(path):
            raise Pg2Error(abi.PG2_ERR_NO_DEVICE, "%s not built (run __graft_entry__.build()); no CPU fallback exists" % path)
        _libs[path] = _bind(C.CDLL(path))
    return _libs[path]


def find_prefix_anchors(seq1, seq2, min_length, lib=None):
    """pg2_find_prefix_anchors on two byte strings -> int32 array (n, 3) of (start_1, start_2, length), the reference's order
    (Find_anchors::find_long_substrings, utils/find_anchors.cpp:35-127).  Host function of the C-ABI: needs no device."""
    lib = lib or load_library()
    cap = 1024
    while True:
        out = np.zeros((cap, 3), np.int32)
        n = C.c_int32()
        rc = lib.pg2_find_prefix_anchors(seq1, len(seq1), seq2, len(seq2), min_length, out.ctypes.data, cap, C.byref(n))
        if rc == abi.PG2_ERR_CAPACITY:
            cap = n.value
            continue
        if rc != abi.PG2_OK:
            raise Pg2Error(rc, "pg2_find_prefix_anchors")
        return out[: n.value].copy()


def anchor_band(hits, str1, str2, width, lib=None):
    """pg2_anchor_band: (n, 3) int32 hits + the two gapped sequence strings -> (upper, lower), len(str1) + 1 values each
    (Find_anchors::define_tunnel, utils/find_anchors.cpp:320-435).  Host function of the C-ABI."""
    lib = lib or load_library()
    hits = np.ascontiguousarray(hits, np.int32).reshape(-1, 3)
    upper = np.zeros(len(str1) + 1, np.int32)
    lower = np.zeros(len(str1) + 1, np.int32)
    rc = lib.pg2_anchor_band(hits.ctypes.data, len(hits), str1, len(str1), str2, len(str2), width, upper.ctypes.data, lower.ctypes.data)
    if rc != abi.PG2_OK:
        raise Pg2Error(rc, "pg2_anchor_band")
    return upper, lower


RESULT_DTYPE = np.dtype([("score", "<f8"), ("cells", "<i8"), ("step_off", "<i8"), ("n_steps", "<i4"),
                         ("status", "<i4"), ("end_ptr", "<u4"), ("kernel", "<i4")])
assert RESULT_DTYPE.itemsize == C.sizeof(abi.Result)


class Batch:
    """A launch batch resident on the device (pg2_batch_*)."""

    def __init__(self, engine, jobs, compact=False):
        self.engine = engine
        self.jobs = jobs
        self.n = len(jobs)
        self._structs = (abi.Job * max(self.n, 1))()
        for k, j in enumerate(jobs):
            self._structs[k] = j.as_struct(engine.model_handle(j.model), compact)
        self._h = C.c_void_p()
        engine._check(engine.lib.pg2_batch_create(engine.ctx, self.n, self._structs, C.byref(self._h)))
        self.step_capacity = int(engine.lib.pg2_batch_step_capacity(self._h))

    def run(self):
        self.engine._check(self.engine.lib.pg2_batch_run(self.engine.ctx, self._h))

    def fetch(self):
        results = np.zeros(max(self.n, 1), dtype=RESULT_DTYPE)
        steps = np.zeros(max(self.step_capacity, 1), dtype=np.uint16)
        self.engine._check(self.engine.lib.pg2_batch_fetch(
            self.engine.ctx, self._h, results.ctypes.data_as(C.POINTER(abi.Result)), steps.ctypes.data, steps.shape[0]))
        return results[: self.n], steps

    def device_buffers(self):
        """pg2_batch_device_buffers -> (address of the 24-byte result records, address of the packed-pointer
        buffer, its length in uint16) of the last run, for GPU-to-GPU forwarding (shard.py)."""
        rp, sp, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        self.engine._check(self.engine.lib.pg2_batch_device_buffers(self.engine.ctx, self._h, C.byref(rp), C.byref(sp), C.byref(n)))
        return rp.value or 0, sp.value or 0, n.value

    def close(self):
        if self._h:
            self.engine.lib.pg2_batch_destroy(self.engine.ctx, self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def results_from_records(records, step_off, jobs):
    """RESULT_DTYPE array (what pg2_batch_fetch returns) from gathered shard records (shard.assemble)."""
    res = np.zeros(len(jobs), dtype=RESULT_DTYPE)
    res["score"] = records["score"]
    res["cells"] = [j.cells for j in jobs]
    res["step_off"] = step_off
    res["n_steps"] = records["n_steps"]
    res["status"] = np.where(records["status"] == 5, abi.PG2_JOB_BAD_GRAPH, records["status"])  # JOB_UNSUPPORTED, as pg2_batch_fetch
    res["end_ptr"] = records["end_ptr"]
    res["kernel"] = -1
    return res


class Engine:
    def __init__(self, device=0, lib_path=None):
        self.lib = load_library(lib_path)
        self.ctx = C.c_void_p()
        self._models = {}
        self._check(self.lib.pg2_ctx_create(device, C.byref(self.ctx)))

    def _check(self, rc):
        if rc != abi.PG2_OK:
            raise Pg2Error(rc, self.lib.pg2_last_error().decode())

    def close(self):
        if self.ctx:
            self.lib.pg2_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def model_handle(self, model):
        key = id(model)
        if key not in self._models:
            h = C.c_int32()
            ms = model.as_struct()
            self._check(self.lib.pg2_model_upload(self.ctx, C.byref(ms), C.byref(h)))
            self._models[key] = (h.value, model)
        return self._models[key][0]

    def batch(self, jobs, compact=False):
        return Batch(self, jobs, compact)

    def prepare(self, jobs, pinned=False, compact=False):
        """Builds the pg2_job array (the caller-side host buffers of the C-ABI) once; the arrays of the
        FlatJob objects stay owned by `jobs`.  pinned=True puts the result / step buffers in page-locked
        host memory (torch), which lets the device->host copy run at full PCIe rate.  compact=True describes plain
        chains (reads, leaves) by their states only (the compact pg2_graph form).
        Returns an opaque tuple for align_prepared()."""
        n = len(jobs)
        structs = (abi.Job * max(n, 1))()
        cap = 0
        for k, j in enumerate(jobs):
            structs[k] = j.as_struct(self.model_handle(j.model), compact)
            cap += j.left.n_sites + j.right.n_sites
        if pinned:
            import torch

            rbuf = torch.zeros(max(n, 1) * RESULT_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
            sbuf = torch.zeros(max(cap, 1), dtype=torch.int16, pin_memory=True)
            results = rbuf.numpy().view(RESULT_DTYPE)
            steps = sbuf.numpy().view(np.uint16)
            return (n, structs, results, steps, (jobs, rbuf, sbuf))
        results = np.zeros(max(n, 1), dtype=RESULT_DTYPE)
        steps = np.zeros(max(cap, 1), dtype=np.uint16)
        return (n, structs, results, steps, jobs)

    def align_prepared(self, prep):
        """One pg2_align_batch call: host job arrays in, host results + packed steps out."""
        n, structs, results, steps, _ = prep
        self._check(self.lib.pg2_align_batch(self.ctx, n, structs, results.ctypes.data_as(C.POINTER(abi.Result)),
                                             steps.ctypes.data, steps.shape[0]))
        return results[:n], steps

    def align(self, jobs):
        """pg2_align_batch on a list of FlatJob -> (results[RESULT_DTYPE], packed steps uint32)."""
        return self.align_prepared(self.prepare(jobs))

    def expand(self, job, result, steps, compact=False):
        """pg2_expand_path -> (steps[STEP_DTYPE] forward order, used_left, used_right)."""
        cap = job.left.n_sites + job.right.n_sites
        out = np.zeros(cap, dtype=abi.STEP_DTYPE)
        ul = np.zeros(cap, np.int32)
        ur = np.zeros(cap, np.int32)
        n, nl, nr = C.c_int32(), C.c_int32(), C.c_int32()
        js = job.as_struct(0, compact)
        ms = job.model.as_struct()
        r = abi.Result()
        for f, _ in abi.Result._fields_:
            setattr(r, f, result[f].item())
        self._check(self.lib.pg2_expand_path(C.byref(js), C.byref(ms), C.byref(r), steps.ctypes.data, out.ctypes.data,
                                             C.byref(n), ul.ctypes.data, C.byref(nl), ur.ctypes.data, C.byref(nr)))
        return out[: n.value].copy(), ul[: nl.value].copy(), ur[: nr.value].copy()

    def stats(self):
        s = abi.Stats()
        self._check(self.lib.pg2_get_stats(self.ctx, C.byref(s)))
        return {f: getattr(s, f) for f, _ in abi.Stats._fields_}
