"""Multi-GPU sharding of one launch batch.

Alignments of a launch batch are independent (a guide-tree wave, node.cpp:240-264; the trial / final
alignments of many reads, reads_aligner.cpp:983-1216), so the batch is cut by index range over the ranks,
every rank runs its shard on its own pg2_ctx, and ONE exchange step brings the fixed-size result records
and the run-length encoded, compacted traceback pointers to rank 0, which is where the reference's serial host step
(build_ancestral_sequence, basic_alignment.cpp:36-59) consumes them.  No collective runs inside an
alignment.  The exchange is torch.distributed (NCCL over NVLink on the GPU box, straight from the device
buffers pg2_batch_device_buffers exposes; gloo in the CPU tests).
"""
import ctypes as C

import numpy as np

RECORD_BYTES = 24  # pg2_device.cuh DevResult {double score; uint32 end_ptr; int32 n_steps; int32 status; int32 pad = raw step count}
RECORD_DTYPE = np.dtype([("score", "<f8"), ("end_ptr", "<u4"), ("n_steps", "<i4"), ("status", "<i4"), ("pad", "<i4")])
assert RECORD_DTYPE.itemsize == RECORD_BYTES


def partition(cells, world, groups=None, unit=32):
    """Deterministic split of job indices over `world` ranks.  Returns one int64 index array per rank (possibly empty).

    groups=None: contiguous index ranges of the batch after a size-interleaving permutation (jobs sorted by cell count,
    largest first, dealt round-robin), so that every range carries the same mix of job sizes.

    groups = one key per job (placement: the target node a read is aligned against): the jobs of a group stay together.
    The placement kernel sweeps 32 reads that share their target per task (pg2_lanes.cu), so a cut through a group costs
    every rank a partly filled task per group -- at 8 ranks and 127 targets that is a quarter of the lanes idle.  Here the
    groups (largest first) are laid end to end in units of `unit` jobs of similar size and the unit sequence is cut into
    `world` contiguous ranges of equal cell count: a rank holds whole groups except at its two borders, so it also uploads
    only the target graphs it needs."""
    cells = np.asarray(cells, dtype=np.int64)
    n = cells.shape[0]
    if groups is None or n == 0:
        order = np.argsort(-cells, kind="stable")
        perm = np.concatenate([order[r::world] for r in range(world)]) if n else order
        bounds = [0]
        for r in range(world):
            bounds.append(bounds[-1] + len(order[r::world]))
        return [perm[bounds[r]:bounds[r + 1]] for r in range(world)]
    keys = {}
    gid = np.empty(n, dtype=np.int64)
    for t, g in enumerate(groups):  # first-appearance ids: the result does not depend on the key values themselves
        gid[t] = keys.setdefault(g, len(keys))
    ng = len(keys)
    gcells = np.bincount(gid, weights=cells, minlength=ng)
    grank = np.empty(ng, dtype=np.int64)
    grank[np.argsort(-gcells, kind="stable")] = np.arange(ng)
    # jobs ordered by (group, largest first inside the group); units never span two groups
    order = np.lexsort((np.arange(n), -cells, grank[gid]))
    og = gid[order]
    starts = np.flatnonzero(np.concatenate([[True], og[1:] != og[:-1]]))
    unit_bounds = []
    for a, b in zip(starts, list(starts[1:]) + [n]):
        unit_bounds.extend(range(a, b, unit))
    unit_bounds.append(n)
    unit_bounds = np.asarray(unit_bounds, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(cells[order])])
    ucum = csum[unit_bounds]  # cells before each unit boundary
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        k = int(np.searchsorted(ucum, total * r / world, side="left"))
        k = min(max(k, 0), len(unit_bounds) - 1)
        if k > 0 and abs(ucum[k - 1] - total * r / world) <= abs(ucum[k] - total * r / world):
            k -= 1
        cuts.append(max(int(unit_bounds[k]), cuts[-1]))
    cuts.append(n)
    return [order[cuts[r]:cuts[r + 1]] for r in range(world)]


GENERAL_VROW_COST = 2.0  # a virtual row of a multi-edge site against a plain row of the placement kernel (measured: the shard that
                         # holds the root-most targets of the bench tree takes 1.4 - 1.7 times longer per cell than the shard of leaves)


def placement_costs(jobs):
    """Relative device time of placement jobs for partition(): DP columns of the read x (plain rows + GENERAL_VROW_COST x
    virtual rows of the target's other sites: one per backward edge).  The cell count alone puts the root-most nodes of the
    reference tree -- a third of their sites carry several edges -- on one rank as if they were leaves."""
    per_graph = {}
    out = np.empty(len(jobs), dtype=np.int64)
    for t, j in enumerate(jobs):
        g = j.left
        w = per_graph.get(id(g))
        if w is None:
            n = g.n_sites
            deg = np.diff(g.off)
            first = g.off[:-1]
            idx = np.arange(n, dtype=np.int64)
            has = deg == 1
            plain = np.zeros(n, dtype=bool)
            e = first[has]
            plain[has] = (g.start[e] == idx[has] - 1) & (g.logw.view(np.uint32)[e] == 0)
            plain[0] = True
            w = float(plain.sum()) + GENERAL_VROW_COST * float(deg[~plain].sum())
            per_graph[id(g)] = w
        out[t] = int(w * max(j.right.n_sites - 1, 1))
    return out


def step_capacity(job):
    """Packed-pointer slots the engine reserves for one job (pg2_engine.cu: left.n_sites + right.n_sites)."""
    return job.left.n_sites + job.right.n_sites


class _CudaBuf:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def buffer_views(batch, device):
    """Zero-copy torch views (uint8) of a run batch's result records and packed-pointer buffer.
    device: a torch.device; 'cpu' is only meaningful with the CPU test build of the library."""
    import torch

    rp, sp, total = batch.device_buffers()
    nrec, nst = batch.n * RECORD_BYTES, int(total) * 2
    if device.type == "cuda":
        rec = torch.as_tensor(_CudaBuf(rp, max(nrec, 1)), device=device)[:nrec]
        st = torch.as_tensor(_CudaBuf(sp, max(nst, 1)), device=device)[:nst]
    else:
        rec = torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * max(nrec, 1)).from_address(rp)))[:nrec]
        st = torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * max(nst, 1)).from_address(sp)))[:nst]
    return rec, st


def gather_to_root(dist, rank, world, records, steps, root=0):
    """The exchange step.  records / steps: this rank's uint8 tensors (any length).  Returns on the root a
    list of (records_r, steps_r) uint8 tensors, one per rank in rank order; None elsewhere."""
    import torch

    dev = records.device
    sizes = torch.tensor([records.numel(), steps.numel()], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [s.cpu().tolist() for s in all_sizes]
    out = []
    for col, mine in ((0, records), (1, steps)):
        width = max(max(s[col] for s in all_sizes), 1)
        send = mine
        if mine.numel() != width:
            send = torch.zeros(width, dtype=torch.uint8, device=dev)
            send[: mine.numel()] = mine
        recv = [torch.empty(width, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == root else None
        dist.gather(send, recv, dst=root)
        out.append(recv)
    if rank != root:
        return None
    return [(out[0][r][: all_sizes[r][0]], out[1][r][: all_sizes[r][1]]) for r in range(world)]


def assemble(parts, shards, jobs):
    """Rank 0: merge the gathered shards back into batch order.  Returns (records[RECORD_DTYPE] in job
    order, step_off[int64] per job into `steps`, steps uint16 = the shards' buffers back to back)."""
    n = len(jobs)
    records = np.zeros(n, dtype=RECORD_DTYPE)
    step_off = np.zeros(n, dtype=np.int64)
    bufs, base = [], 0
    for (rec, st), idx in zip(parts, shards):
        rec = rec.cpu().numpy().view(RECORD_DTYPE)
        st = st.cpu().numpy().view(np.uint16)
        assert rec.shape[0] == len(idx)
        # the shard's path words lie back to back in shard-job order (the engine compacts them on the device)
        words = rec["n_steps"].astype(np.int64)
        offs = np.concatenate([[0], np.cumsum(words)[:-1]]) if len(idx) else np.zeros(0, np.int64)
        assert st.shape[0] == int(words.sum())
        records[idx] = rec
        step_off[idx] = base + offs
        bufs.append(st)
        base += st.shape[0]
    steps = np.concatenate(bufs) if bufs else np.zeros(0, np.uint16)
    return records, step_off, steps


def placement_shards(jobs, world):
    """The cut of a placement batch: a target's reads stay together, the ranks get equal estimated device time."""
    return partition(placement_costs(jobs), world, groups=[id(j.left) for j in jobs])


def align_sharded(eng, jobs, dist, rank, world, device, root=0, by_target=False):
    """Runs this rank's shard of `jobs` (every rank holds the same job list) and gathers to the root.
    Returns (records, step_off, steps) on the root, None elsewhere.  by_target: placement_shards instead of the
    size-interleaved cut."""
    shards = placement_shards(jobs, world) if by_target else partition([j.cells for j in jobs], world)
    mine = [jobs[i] for i in shards[rank]]
    batch = eng.batch(mine)
    try:
        batch.run()
        rec, st = buffer_views(batch, device)
        parts = gather_to_root(dist, rank, world, rec, st, root)
        if device.type == "cuda":
            # rec / st are views of the engine's own buffers and the gather is asynchronous (NCCL stream): every rank
            # waits for its sends before the batch is closed and the next batch may rewrite or free those buffers
            import torch

            torch.cuda.current_stream(device).synchronize()
    finally:
        batch.close()
    if rank != root:
        return None
    return assemble(parts, shards, jobs)
