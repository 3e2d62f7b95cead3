"""Seeded random flat jobs for property tests: multi-edge graphs with long-span edges, non-unit
weights, tie-prone parameter sets, random monotone bands."""
import numpy as np

from pagan2_msa_b200.abi import FlatGraph, FlatJob, Model


def random_graph(rng, n_real, fas, max_extra=3, max_span=12, p_extra=0.25, tie_weights=False):
    """n_real real sites + start + stop.  Every site s>=1 has the chain edge (s-1 -> s) somewhere in its
    list plus up to max_extra extra backward edges to earlier sites (span <= max_span)."""
    n = n_real + 2
    state = np.full(n, -1, np.int32)
    state[1:-1] = rng.integers(0, fas, size=n_real)
    off = [0]
    start, logw, eidx = [], [], []
    next_e = 1  # edge 0 is the reference's dummy first edge
    for s in range(n):
        if s > 0:
            preds = [s - 1]
            if rng.random() < p_extra:
                k = int(rng.integers(1, max_extra + 1))
                lo = max(0, s - max_span)
                cand = [p for p in range(lo, s - 1)]
                rng.shuffle(cand)
                preds += cand[:k]
            rng.shuffle(preds)
            for p in preds:
                start.append(p)
                if tie_weights:
                    w = [1.0, 1.0, 0.5, 0.25][int(rng.integers(0, 4))]
                else:
                    w = 1.0 if rng.random() < 0.5 else float(rng.uniform(0.05, 1.0))
                logw.append(np.log(np.float32(w)))
                eidx.append(next_e)
                next_e += 1 + int(rng.random() < 0.1)  # holes: deleted edges stay in the reference's vector
        off.append(len(start))
    return FlatGraph(state, off, start, np.array(logw, np.float32), eidx)


def random_model(rng, fas, ties=False):
    if ties:
        vals = np.array([-2.0, -1.0, 0.5, 1.0], np.float32)
        table = vals[rng.integers(0, len(vals), size=fas * fas)]
        scal = np.array([-3.0, -0.5, -0.25, -0.125, -0.0625], np.float32)
    else:
        table = rng.normal(-1.0, 1.5, size=fas * fas).astype(np.float32)
        d = np.arange(fas)
        t2 = table.reshape(fas, fas)
        t2[d, d] = np.abs(t2[d, d]) + 0.5
        scal = np.array([-rng.uniform(2, 5), -rng.uniform(0.2, 1.0), -rng.uniform(0.05, 0.5), -rng.uniform(0.1, 0.9),
                         -rng.uniform(0.001, 0.1)], np.float32)
    return Model(fas, table, scal)


def random_band(rng, lx, ly, min_w=3, max_w=12):
    """Monotone band around the main diagonal containing (0,0) and (lx-1, ly-1), every row connected
    to the next."""
    upper = np.zeros(lx, np.int64)
    lower = np.zeros(lx, np.int64)
    slope = int(np.ceil((ly - 1) / max(lx - 1, 1)))
    min_w, max_w = min_w + slope, max_w + slope  # consecutive rows must overlap or no path exists
    for i in range(lx):
        c = int(round(i * (ly - 1) / max(lx - 1, 1)))
        upper[i] = c - int(rng.integers(min_w, max_w + 1))
        lower[i] = c + int(rng.integers(min_w, max_w + 1))
    upper = np.maximum.accumulate(upper)
    lower = np.maximum.accumulate(lower)
    upper[0] = min(upper[0], 0)
    lower[-1] = max(lower[-1], ly + 3)  # the reference's bounds may exceed the matrix; clipping is the callee's job
    return upper.astype(np.int32), lower.astype(np.int32)


def random_job(rng, kind="general", fas=15):
    ties = rng.random() < 0.4
    model = random_model(rng, fas, ties)
    flags = int(rng.integers(0, 4))
    if kind == "strip":  # general left graph, plain-chain right graph (placement shape)
        nl, nr = int(rng.integers(3, 90)), int(rng.integers(1, 200))
        left = random_graph(rng, nl, fas, tie_weights=ties)
        right = FlatGraph.chain(rng.integers(0, fas, size=nr).astype(np.int32))
        if rng.random() < 0.3:  # non-unit weights on the chain
            right.logw[:] = np.log(rng.uniform(0.2, 1.0, size=right.logw.shape[0]).astype(np.float32))
        return FlatJob(left, right, model, flags)
    nl, nr = int(rng.integers(1, 70)), int(rng.integers(1, 70))
    if kind == "banded_chain":  # plain chains on both sides inside a band (anchored leaf x leaf): no score scratch reads
        left = FlatGraph.chain(rng.integers(0, fas, size=nl).astype(np.int32))
        right = FlatGraph.chain(rng.integers(0, fas, size=nr).astype(np.int32))
    else:
        left = random_graph(rng, nl, fas, tie_weights=ties)
        right = random_graph(rng, nr, fas, tie_weights=ties)
    job = FlatJob(left, right, model, flags)
    if kind in ("banded", "banded_chain"):
        up, lo = random_band(rng, left.n_sites - 1, right.n_sites - 1)
        job.upper, job.lower = up, lo
    return job


def random_shared_target_jobs(rng, n_jobs, fas=15, plain_left=False, weights=None, nl=None, nr_max=70):
    """Placement-shaped launch batch: n_jobs reads (plain chains of varying length) against ONE left graph,
    same model and flags -- what the lane-per-alignment kernel groups into tasks of 32."""
    ties = rng.random() < 0.4
    model = random_model(rng, fas, ties)
    flags = int(rng.integers(0, 4))
    nl = int(rng.integers(3, 90)) if nl is None else nl
    if plain_left:
        left = FlatGraph.chain(rng.integers(0, fas, size=nl).astype(np.int32))
    else:
        left = random_graph(rng, nl, fas, tie_weights=ties)
    if weights is None:
        weights = rng.random() < 0.3
    jobs = []
    for _ in range(n_jobs):
        nr = int(rng.integers(1, nr_max))
        right = FlatGraph.chain(rng.integers(0, fas, size=nr).astype(np.int32))
        if weights and rng.random() < 0.5:
            right.logw[:] = np.log(rng.uniform(0.2, 1.0, size=right.logw.shape[0]).astype(np.float32))
        jobs.append(FlatJob(left, right, model, flags))
    return jobs


def random_anchor_band_job(rng, n, fas=15, width=(3, 12), bulge=0, p_indel=0.02, disconnect=False):
    """Anchored leaf x leaf shape: two plain unit-weight chains of ~n sites (the second a copy with substitutions and
    indels), a monotone band of varying half-width around the true path; `bulge` adds stretches where the band is that
    much wider (gaps between anchors), `disconnect` cuts the band in two (no path: PG2_JOB_NO_PATH)."""
    a = rng.integers(0, min(fas, 4), size=n).astype(np.int32)
    b, col = [], []
    for x in a:
        r = rng.random()
        if r < p_indel:  # deletion in b
            col.append(max(len(b) - 1, 0))
            continue
        if r < 2 * p_indel:  # insertion in b
            b.extend(rng.integers(0, min(fas, 4), size=int(rng.integers(1, 6))).tolist())
        col.append(len(b))
        b.append(int(x) if rng.random() > 0.05 else int(rng.integers(0, min(fas, 4))))
    if not b:
        b = [0]
    left, right = FlatGraph.chain(a), FlatGraph.chain(np.array(b, np.int32))
    lx, ly = left.n_sites - 1, right.n_sites - 1
    c = np.concatenate([[0], np.array(col, np.int64) + 1])[:lx]
    w_lo = rng.integers(width[0], width[1] + 1, size=lx)
    w_hi = rng.integers(width[0], width[1] + 1, size=lx)
    if bulge:
        for _ in range(max(1, lx // 300)):
            s0 = int(rng.integers(0, lx))
            w_hi[s0:s0 + int(rng.integers(5, 60))] += int(rng.integers(1, bulge + 1))
            s0 = int(rng.integers(0, lx))
            w_lo[s0:s0 + int(rng.integers(5, 60))] += int(rng.integers(1, bulge + 1))
    upper = np.maximum.accumulate(c - w_lo)
    lower = np.maximum.accumulate(c + w_hi)
    upper[0] = min(upper[0], 0)
    lower[-1] = max(lower[-1], ly + 3)
    job = FlatJob(left, right, random_model(rng, fas, rng.random() < 0.3), int(rng.integers(0, 4)))
    if disconnect and lx > 8:
        k = lx // 2
        upper[k:] = np.maximum(upper[k:], lower[k - 1] + 2)  # rows k.. start right of where row k-1 ends
        lower[k:] = np.maximum(lower[k:], upper[k:] + 1)
        lower = np.maximum.accumulate(lower)
    job.upper, job.lower = upper.astype(np.int32), lower.astype(np.int32)
    return job
