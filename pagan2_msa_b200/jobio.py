"""Reader/writer for PJOB job streams.

A job stream is what oracle/_ref (the reference compiled with the dump interposer, oracle/ref_glue.cpp)
writes for every Viterbi_alignment::align call: inputs in the flat CSR layout of include/pagan2_b200.h
plus the reference's score and path.  The same container is used for the committed fixtures under
tests/golden/.

Container (little endian): file = records; record = u32 'PJOB', u32 n_fields;
field = u32 name_len, name, u32 dtype (0=i32, 1=f32, 2=f64), u64 count, payload.

Large fixtures (200 kb graphs, 400 000-step paths) are stored compactly by save_jobs(..., compact=True): an int32
field "name" may be stored as "name~d" (first differences, which gzip then squeezes: CSR offsets, edge starts, edge
indices and path columns are near-arithmetic), and the per-step scores of a long path as "path_score_sha" (the SHA-256
of their bytes) instead of the 8 bytes per step.
"""
import gzip
import hashlib
import struct

import numpy as np

from .abi import FlatGraph, FlatJob, Model

_DT = {0: np.dtype("<i4"), 1: np.dtype("<f4"), 2: np.dtype("<f8")}
_CODE = {np.dtype("<i4"): 0, np.dtype("<f4"): 1, np.dtype("<f8"): 2}
_MAGIC = 0x424F4A50


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def read_records(path):
    with _open(path, "rb") as f:
        buf = f.read()
    pos, out = 0, []
    while pos < len(buf):
        magic, nf = struct.unpack_from("<II", buf, pos)
        if magic != _MAGIC:
            raise ValueError("bad PJOB magic at offset %d" % pos)
        pos += 8
        rec = {}
        for _ in range(nf):
            (nl,) = struct.unpack_from("<I", buf, pos)
            pos += 4
            name = buf[pos : pos + nl].decode()
            pos += nl
            dt, cnt = struct.unpack_from("<IQ", buf, pos)
            pos += 12
            d = _DT[dt]
            rec[name] = np.frombuffer(buf, dtype=d, count=cnt, offset=pos).copy()
            pos += cnt * d.itemsize
        out.append(rec)
    return out


def write_records(path, records):
    with _open(path, "wb") as f:
        for rec in records:
            f.write(struct.pack("<II", _MAGIC, len(rec)))
            for name, arr in rec.items():
                a = np.ascontiguousarray(arr)
                if a.dtype not in _CODE:
                    raise ValueError("unsupported dtype %s for field %s" % (a.dtype, name))
                nb = name.encode()
                f.write(struct.pack("<I", len(nb)))
                f.write(nb)
                f.write(struct.pack("<IQ", _CODE[a.dtype], a.size))
                f.write(a.tobytes())


COMPACT_MIN = 4096      # int32 fields at least this long are delta-coded in compact fixtures
SCORE_SHA_MIN = 20000   # paths at least this long keep a digest of their per-step scores instead of the scores


def sha_words(arr):
    """SHA-256 of an array's bytes as 8 int32 words (what 'path_score_sha' holds)."""
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(arr).tobytes()).digest(), dtype="<i4").copy()


def _undelta(rec):
    out = {}
    for name, arr in rec.items():
        if name.endswith("~d"):
            out[name[:-2]] = np.cumsum(arr.astype(np.int64)).astype("<i4")
        else:
            out[name] = arr
    return out


def _delta(rec):
    out = {}
    for name, arr in rec.items():
        a = np.ascontiguousarray(arr)
        if a.dtype == np.dtype("<i4") and a.size >= COMPACT_MIN and name != "path":
            d = a.astype(np.int64)
            d[1:] = d[1:] - d[:-1]
            if np.abs(d).max() < 2 ** 31:
                out[name + "~d"] = d.astype("<i4")
                continue
        out[name] = a
    return out


def records_to_jobs(records):
    """PJOB records -> list[FlatJob] with expected results attached; tables are shared by table_id."""
    tables = {}
    jobs = []
    for rec in records:
        rec = _undelta(rec)
        if "path_cols" in rec:  # compact form: the six path columns one after the other
            rec["path"] = np.ascontiguousarray(rec["path_cols"].reshape(6, -1).T).reshape(-1)
        meta = rec["meta"]
        fas, flags, table_id = int(meta[0]), int(meta[1]), int(meta[7])
        if "table" in rec:
            tables[table_id] = rec["table"]
        model = Model(fas, tables[table_id], rec["model"])
        left = FlatGraph(rec["l_state"], rec["l_off"], rec["l_start"], rec["l_logw"], rec["l_eidx"])
        right = FlatGraph(rec["r_state"], rec["r_off"], rec["r_start"], rec["r_logw"], rec["r_eidx"])
        job = FlatJob(left, right, model, flags, rec.get("upper"), rec.get("lower"))
        job.expected_score = float(rec["score"][0])
        job.expected_path = rec["path"].reshape(-1, 6)
        job.expected_path_score = rec.get("path_score")
        if "path_score_sha" in rec:
            job.expected_path_score_sha = rec["path_score_sha"]
        # is_used edge marks around the reference's call (newer dumps): {side: (before, after)} index arrays
        if "l_used_after" in rec:
            job.expected_used = {"l": (rec["l_used_before"], rec["l_used_after"]), "r": (rec["r_used_before"], rec["r_used_after"])}
        job.meta = {
            "data_type": int(meta[2]),
            "is_reads": int(meta[3]),
            "banded": int(meta[4]),
            "table_id": table_id,
            "dist": rec["dist"].tolist(),
            "ref_time": rec["time"].tolist() if "time" in rec else None,
        }
        jobs.append(job)
    return jobs


def jobs_to_records(jobs, compact=False):
    """Inverse of records_to_jobs (tables de-duplicated by content)."""
    seen = {}
    out = []
    for job in jobs:
        key = job.model.table.tobytes()
        new = key not in seen
        if new:
            seen[key] = len(seen)
        tid = seen[key]
        cells = job.cells
        rec = {
            "meta": np.array(
                [job.model.fas, job.flags, job.meta.get("data_type", 0), job.meta.get("is_reads", 0),
                 0 if job.upper is None else 1, job.left.n_sites, job.right.n_sites, tid,
                 cells & 0x7FFFFFFF, cells >> 31], dtype="<i4"),
            "model": job.model.scalars.astype("<f4"),
            "dist": np.array(job.meta.get("dist", [0, 0, 0]), dtype="<f4"),
        }
        if new:
            rec["table"] = job.model.table.astype("<f4")
        for side, g in (("l", job.left), ("r", job.right)):
            rec[side + "_state"] = g.state
            rec[side + "_off"] = g.off
            rec[side + "_start"] = g.start
            rec[side + "_logw"] = g.logw
            rec[side + "_eidx"] = g.eidx
        if job.upper is not None:
            rec["upper"] = job.upper
            rec["lower"] = job.lower
        rec["score"] = np.array([job.expected_score], dtype="<f8")
        path = np.ascontiguousarray(job.expected_path, dtype="<i4").reshape(-1, 6)
        ps = job.expected_path_score
        if compact and path.shape[0] >= COMPACT_MIN:
            rec["path_cols"] = np.ascontiguousarray(path.T).reshape(-1)  # column-major: delta-codes well
        else:
            rec["path"] = path.reshape(-1)
        if ps is None:
            rec["path_score_sha"] = np.ascontiguousarray(job.expected_path_score_sha, dtype="<i4")
        elif compact and path.shape[0] >= SCORE_SHA_MIN:
            rec["path_score_sha"] = sha_words(np.ascontiguousarray(ps, dtype="<f8"))
        else:
            rec["path_score"] = np.ascontiguousarray(ps, dtype="<f8")
        used = getattr(job, "expected_used", None)
        if used is not None:
            for side in ("l", "r"):
                rec[side + "_used_before"] = np.ascontiguousarray(used[side][0], dtype="<i4")
                rec[side + "_used_after"] = np.ascontiguousarray(used[side][1], dtype="<i4")
        out.append(_delta(rec) if compact else rec)
    return out


def load_jobs(path):
    return records_to_jobs(read_records(path))


def save_jobs(path, jobs, compact=False):
    write_records(path, jobs_to_records(jobs, compact))
