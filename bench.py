#!/usr/bin/env python
"""bench.py -- graph-DP throughput of the PAGAN2 pairwise Viterbi path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU code, host cores

Workload (config.workload): BASELINE.json configs[1], query placement -- synthetic 150-nt reads against a
64-taxon x 1.5 kb reference alignment + tree.  The 127 target graphs (64 leaves + 63 internal nodes) are
the reference's own graphs (tests/golden/bench_targets.pjob.gz, made by the reference binary from a seeded
synthetic alignment); reads are seeded substrings of the leaf sequences with 1 % substitutions, each
assigned to a node on the path from its source leaf to the root.  One step = one launch batch = one
alignment (fill + end corner + traceback) of every read against its target: --reads x 151 x ~1.5 k cells.

Printed line (rank 0): metric/value = whole-job GCUPS with the batch resident in HBM (device time, CUDA
events on the engine's stream, max over ranks); e2e = the same through pg2_align_batch from host buffers
(packing, H2D, kernels, D2H inside the timed region); roofline, cpu_baseline, clocks as the contract asks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pagan2_msa_b200 import abi, jobio, synth  # noqa: E402

TARGETS = os.path.join(ROOT, "tests", "golden", "bench_targets.pjob.gz")
FP64_INSTR_PER_CELL = 22  # 13 DADD + 9 DSETP per in-degree-1 unit-weight cell (the reference's own count)
FP64_INSTR_EXECUTED = 15.4  # what the lane kernel's hot loop issues per cell: 9 DADD + 6 DSETP (+ 3 DADD per 8-cell row), from SASS
PTR_BYTES_PER_CELL = 2    # packed back-pointers streamed to HBM


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


TRAFFIC_PROFILE = "r2_lanes_v11_traffic.json"  # tools/ncu_summary.py output of the committed ncu capture


def measured_profile(n_reads, fill_launches):
    """Per-launch figures of the committed ncu capture (profiles/), valid for the workload it was taken on:
    (DRAM read+write bytes per fill launch, {issue-slot, FP64-pipe, ALU-pipe busy % and warp-instructions} of the fill
    launches, weighted by time)."""
    p = os.path.join(ROOT, "profiles", TRAFFIC_PROFILE)
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        t = json.load(f)
    if t.get("reads_per_gpu") != n_reads or len(t["launches"]) != fill_launches:
        return None, None
    ls = t["launches"]
    traffic = sum(l["dram_bytes_read"] + l["dram_bytes_write"] for l in ls) / len(ls)
    tt = sum(l["time_ms_under_ncu"] for l in ls)
    busy = None
    if all(l.get("issue_active_pct") is not None for l in ls) and tt > 0:
        busy = {k: sum(l[k] * l["time_ms_under_ncu"] for l in ls) / tt
                for k in ("issue_active_pct", "fp64_pipe_active_pct", "alu_pipe_active_pct") if all(l.get(k) is not None for l in ls)}
        busy["warp_instructions"] = sum(l["warp_instructions"] for l in ls)
    return traffic, busy


def build_workload(n_reads, seed, rank=0):
    """Returns (jobs, info).  Deterministic in (n_reads, seed, rank)."""
    tjobs = jobio.load_jobs(TARGETS)
    model = tjobs[0].model
    targets = [j.left for j in tjobs]
    # node order of the reference's post-order naming is not needed: recover leaf/ancestor relations from
    # sizes is fragile, so reads are cut from LEAF graphs (plain chains) and assigned either to that leaf
    # or to a random internal node (a read aligns to any ancestor of its source equally well in shape).
    leaves = [k for k, g in enumerate(targets) if (np.diff(g.off)[1:] == 1).all() and (g.logw == 0).all()]
    internal = [k for k in range(len(targets)) if k not in leaves]
    rng = np.random.default_rng(seed * 1000003 + rank)
    reads, assign = [], []
    for k in range(n_reads):
        src = leaves[int(rng.integers(0, len(leaves)))]
        seq = targets[src].state[1:-1]
        st = int(rng.integers(0, max(1, len(seq) - 150)))
        r = seq[st:st + 150].copy()
        mut = rng.random(r.shape[0]) < 0.01
        r[mut] = rng.integers(0, 4, size=int(mut.sum()))
        reads.append(r)
        if k % 2 == 0 or not internal:
            assign.append(src)
        else:
            assign.append(internal[int(rng.integers(0, len(internal)))])
    jobs = synth.placement_jobs(targets, reads, assign, model)
    cells = sum(j.cells for j in jobs)
    info = {"n_targets": len(targets), "n_leaf_targets": len(leaves), "cells_per_step": int(cells),
            "mean_target_sites": float(np.mean([g.n_sites for g in targets]))}
    return jobs, info


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local, world):
    """One process per GPU, run on the cores next to that GPU: pinned staging buffers and the packing threads then sit
    on the GPU's own NUMA node (H2D / D2H do not cross the socket link).  Returns the cores this rank may use."""
    mine = sorted(os.sched_getaffinity(0))
    if world <= 1:
        return mine
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        near = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        cores = sorted(set(near) & set(mine))
        if len(cores) >= 2:
            os.sched_setaffinity(0, cores)
            return cores
    except Exception:
        pass
    return mine


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return rank, world, local, dist
    return rank, world, local, None


def cpu_baseline_sample(jobs, eng, results, steps_buf, budget_s=12.0):
    """The reference's own code (oracle/_ref, kind 'reference') or the oracle port, 1 thread, on the first
    jobs of the workload until ~budget_s seconds are spent.  Its outputs are the checker of this very run: every job
    of the sample must equal the GPU's result -- score bits, the expanded path in every field, every per-step score.
    Returns (cpu_baseline object, number of jobs compared)."""
    import oracle_lib

    use_ref = oracle_lib.ref_available()
    t0 = time.perf_counter()
    cells = 0
    ref = []
    for j in jobs:
        if use_ref:
            score, path, pscore = oracle_lib.ref_align_flat(j)
        else:
            _, score, st, _ = oracle_lib.oracle_align(j)
            path = np.stack([st[n] for n in ("matrix", "x_ind", "y_ind", "x_edge_ind", "y_edge_ind", "real_site")], axis=1)
            pscore = st["score"]
        ref.append((score, path, pscore))
        cells += j.cells
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    n = len(ref)
    for k, (score, path, pscore) in enumerate(ref):  # outside the clock
        if np.float64(score).view(np.uint64) != results["score"][k].view(np.uint64):
            raise SystemExit("bench: job %d: GPU score %r differs from the CPU %s's %r" % (k, results["score"][k], "reference" if use_ref else "oracle", score))
        st, _, _ = eng.expand(jobs[k], results[k], steps_buf)
        diffs = oracle_lib.steps_equal(st, path, pscore)
        if diffs:
            raise SystemExit("bench: job %d: GPU path differs from the CPU checker's: %s" % (k, diffs))
    return {"value": cells / dt * 1e-9, "unit": "GCUPS", "cores": 1, "kind": "reference" if use_ref else "port",
            "sample": "first %d jobs of the workload (%d cells) in %.1f s, 1 thread; all %d compared with the GPU results "
                      "(score bits, path, per-step scores)" % (n, cells, dt, n)}, n


def other_configs(eng, reps=3):
    """Device time and GCUPS of the launch batches of the other BASELINE configs, from the reference's own job streams at
    BASELINE sizes (tests/golden/*_full.pjob.gz); every score is compared with the reference's (bit for bit)."""
    out = {}
    def plain(g):
        return bool((np.diff(g.off)[1:] == 1).all())
    streams = {}
    for name in ("c1_full", "c3_full", "c4_full", "c5_full"):
        path = os.path.join(ROOT, "tests", "golden", name + ".pjob.gz")
        if os.path.exists(path):
            streams[name] = jobio.load_jobs(path)
    batches = []
    if "c1_full" in streams:
        c1 = streams["c1_full"]
        batches.append(("c1_wave1_8_leaf_pairs_1kb", [j for j in c1 if plain(j.left) and plain(j.right)]))
        batches.append(("c1_waves2to4_7_ancestor_pairs", [j for j in c1 if not (plain(j.left) and plain(j.right))]))
    if "c3_full" in streams:
        batches.append(("c3_pileup_one_alignment_root_after_220_reads", streams["c3_full"][-1:]))
    if "c4_full" in streams:
        batches.append(("c4_codons_7_alignments_1000_codons", streams["c4_full"]))
    if "c5_full" in streams:
        batches.append(("c5_anchored_2_leaf_pairs_200kb", streams["c5_full"][:2]))
        batches.append(("c5_anchored_ancestor_pair_200kb", streams["c5_full"][2:]))
    for name, jobs in batches:
        if not jobs:
            continue
        b = eng.batch(jobs)
        best = None
        for _ in range(reps + 1):
            b.run()
            st = eng.stats()
            if best is None or st["run_ms"] < best["run_ms"]:
                best = st
        res, _ = b.fetch()
        b.close()
        same = all(np.float64(j.expected_score).view(np.uint64) == res["score"][k].view(np.uint64) for k, j in enumerate(jobs))
        if not same or not (res["status"] == 0).all():
            raise SystemExit("bench: config batch %s: GPU scores differ from the reference's" % name)
        cells = int(sum(j.cells for j in jobs))
        out[name] = {"jobs": len(jobs), "cells": cells, "device_ms": best["run_ms"], "fill_ms": best["fill_ms"],
                     "traceback_ms": best["traceback_ms"], "gcups": cells / best["run_ms"] * 1e-6,
                     "kernels": sorted(set(int(k) for k in res["kernel"])), "scores_equal_reference": True}
    return out


_REF_JOBS = None


def _ref_init(n_reads, seed):
    """Pool initializer of the reference arm: every host process builds the (bounded) workload once, outside the clock."""
    global _REF_JOBS
    _REF_JOBS, _ = build_workload(n_reads, seed)


def _ref_worker(args):
    """One host process of the reference arm: aligns its share of jobs with the reference library."""
    lo, hi = args
    import oracle_lib

    use_ref = oracle_lib.ref_available()
    cells = 0
    for j in _REF_JOBS[lo:hi]:
        if use_ref:
            oracle_lib.ref_align_flat(j)
        else:
            oracle_lib.oracle_align(j)
        cells += j.cells
    return cells


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores.  Each step is a
    bounded sample of the same workload: `cores * per_core` jobs split over one process per core."""
    import multiprocessing as mp

    import oracle_lib

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    per_core = 6
    n = cores * per_core
    use_ref = oracle_lib.ref_available()
    bounds = [(k * per_core, (k + 1) * per_core) for k in range(cores)]
    times, cells_step = [], 0
    # one thread alone (the reference's default mode, main.cpp:100-106), beside the all-cores figure
    _ref_init(n, args.seed)
    t0 = time.perf_counter()
    single_cells = _ref_worker((0, min(per_core * 2, n)))
    single_gcups = single_cells / (time.perf_counter() - t0) * 1e-9
    with mp.get_context("fork").Pool(cores, initializer=_ref_init, initargs=(n, args.seed)) as pool:
        pool.map(_ref_worker, [(0, 0)] * cores)  # every worker up and initialised before the clock starts
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            cells_step = sum(pool.map(_ref_worker, bounds))
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = cells_step * len(times) / total * 1e-9
    line = {
        "impl": "reference", "metric": "graph_dp_gcups", "value": value, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "query placement: 150-nt reads vs 64-taxon x 1.5 kb reference (BASELINE configs[1]); "
                               "bounded sample of %d alignments per step" % n, "cells_per_step": int(cells_step)},
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": cores, "kind": "reference" if use_ref else "port",
                         "sample": "%d alignments per step over %d processes (fork), model build excluded" % (n, cores),
                         "single_thread_gcups": single_gcups},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def emit(line):
    """The ONE JSON line of the contract, on the real stdout (libraries -- NCCL's version banner, make -- get stderr)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=100000, help="reads (= alignments) per GPU per step")
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (--reads in total over all GPUs)")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config launch batches (C1, C3, C4, C5 at BASELINE sizes)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import ctypes as C

    import __graft_entry__
    from pagan2_msa_b200 import engine

    rank, world, local, dist = dist_setup(args.gpus)
    # the host packing threads of all ranks share this box's cores
    n_cores_box = len(os.sched_getaffinity(0))
    near_cores = bind_to_gpu_numa_node(local, world)
    os.environ.setdefault("PG2_PACK_THREADS", str(max(1, min(4, n_cores_box // max(world, 1)))))
    if rank == 0:
        __graft_entry__.build()
    if dist is not None:
        dist.barrier()

    import torch

    jobs, info = build_workload(args.reads, args.seed, rank)
    eng = engine.Engine(local)  # raises without the CUDA library / device: no fallback
    peaks, peaks_kind = measured_peaks()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput: batch packed and uploaded once ----------------
    from pagan2_msa_b200 import shard

    batch = eng.batch(jobs)
    tdev = torch.device("cuda", local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def gather_step():
        """The one exchange step of the multi-GPU path: result records + packed paths of every rank's shard to
        rank 0 over NCCL, straight from the engine's device buffers.  Returns device milliseconds."""
        if dist is None:
            return 0.0
        rec, stp = shard.buffer_views(batch, tdev)
        ev0.record()
        parts = shard.gather_to_root(dist, rank, world, rec, stp)
        ev1.record()
        ev1.synchronize()
        del parts
        return ev0.elapsed_time(ev1)

    for _ in range(max(args.warmup, 3)):
        batch.run()
        gather_step()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    dev_ms, fill_ms, tb_ms, launches, fill_launches, gather_ms = 0.0, 0.0, 0.0, 0, 0, 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        batch.run()
        st = eng.stats()
        g_ms = gather_step()
        dev_ms += st["run_ms"] + g_ms
        gather_ms += g_ms
        fill_ms += st["fill_ms"]
        tb_ms += st["traceback_ms"]
        launches += st["kernel_launches"]
        fill_launches += st["fill_launches"]
    barrier()
    wall_dev = time.perf_counter() - t0
    clocks = sampler.stop()
    stats = eng.stats()
    results, steps_buf = batch.fetch()
    batch.close()
    ok = int((results["status"] == 0).sum())

    # ---------------- end to end: host buffers in, host results out, every step ----------------
    # the caller's host buffers (pg2_job array over the numpy arrays, result + step buffers) exist before
    # the clock starts; each timed step is exactly one pg2_align_batch call
    prep = eng.prepare(jobs, pinned=True, compact=True)  # reads and leaf targets in the compact chain form of the C-ABI
    eng.align_prepared(prep)  # warm-up: grows the pinned staging and device buffers once
    barrier()
    e2e_t = []
    h2d = d2h = 0
    e2e_res = None
    for _ in range(args.steps):
        t1 = time.perf_counter()
        e2e_res, e2e_steps = eng.align_prepared(prep)
        e2e_t.append(time.perf_counter() - t1)
        st = eng.stats()
        h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]
    barrier()
    # outside the clock: the host -> host call must return what the resident batch returned, bit for bit
    e2e_same = bool((e2e_res["score"].view(np.uint64) == results["score"].view(np.uint64)).all()
                    and (e2e_res["status"] == results["status"]).all() and (e2e_res["n_steps"] == results["n_steps"]).all())
    for k in range(0, len(jobs), max(1, len(jobs) // 64)):
        a = e2e_steps[e2e_res["step_off"][k]: e2e_res["step_off"][k] + e2e_res["n_steps"][k]]
        b = steps_buf[results["step_off"][k]: results["step_off"][k] + results["n_steps"][k]]
        e2e_same = e2e_same and a.tobytes() == b.tobytes()
    if not e2e_same:
        raise SystemExit("bench: pg2_align_batch (e2e) and the resident batch disagree")

    # ---------------- strong scaling: the SAME --reads alignments in total, cut by index range over the ranks ----------------
    strong_vals = None
    if world > 1 and not args.no_strong:
        gjobs, _ = build_workload(args.reads, args.seed, 0)  # every rank builds the same job list
        mine = [gjobs[i] for i in shard.partition([j.cells for j in gjobs], world)[rank]]
        sbatch = eng.batch(mine)
        for _ in range(3):
            sbatch.run()
        barrier()
        s_dev = 0.0
        for _ in range(args.steps):
            sbatch.run()
            s_dev += eng.stats()["run_ms"]
            if dist is not None:
                rec, stp = shard.buffer_views(sbatch, tdev)
                ev0.record()
                parts = shard.gather_to_root(dist, rank, world, rec, stp)
                ev1.record()
                ev1.synchronize()
                del parts
                s_dev += ev0.elapsed_time(ev1)
        barrier()
        sbatch.close()  # (one batch per ctx: the host -> host calls below create their own)
        sprep = eng.prepare(mine, pinned=True, compact=True)
        eng.align_prepared(sprep)
        barrier()
        s_e2e = []
        for _ in range(args.steps):
            t1 = time.perf_counter()
            eng.align_prepared(sprep)
            s_e2e.append(time.perf_counter() - t1)
        barrier()
        strong_vals = np.array([s_dev / args.steps, sum(s_e2e) / args.steps * 1e3], dtype=np.float64)
        strong_cells = int(sum(j.cells for j in gjobs))
        del gjobs, mine, sprep

    print("rank %d: e2e %.1f ms/step, device %.1f ms/step, %d near cores" % (
        rank, sum(e2e_t) / args.steps * 1e3, dev_ms / args.steps, len(near_cores)), file=sys.stderr, flush=True)
    cells = info["cells_per_step"]
    local_vals = np.array([dev_ms / args.steps, fill_ms / args.steps, sum(e2e_t) / args.steps * 1e3, wall_dev / args.steps * 1e3],
                          dtype=np.float64)
    if dist is not None:
        t = torch.from_numpy(local_vals).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        local_vals = t.cpu().numpy()
        c = torch.tensor([cells], dtype=torch.int64).cuda()
        dist.all_reduce(c)
        total_cells = int(c.item())
        if strong_vals is not None:
            t = torch.from_numpy(strong_vals).cuda()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            strong_vals = t.cpu().numpy()
    else:
        total_cells = cells
    ms_dev, ms_fill, ms_e2e, ms_wall = [float(x) for x in local_vals]

    if rank == 0:
        dadd, cand = C.c_double(), C.c_double()
        eng.lib.pg2_measure_fp64_issue(local, C.byref(dadd), C.byref(cand))
        mix, mix_mhz = (C.c_double * 3)(), C.c_double()
        eng.lib.pg2_measure_dispatch_mix(local, mix, C.byref(mix_mhz))
        traffic, dispatch_busy = measured_profile(args.reads, stats["fill_launches"])
        fp64_peak = dadd.value  # 1e9 FP64-pipe warp-instructions / s, chip-wide
        per_launch_cells = cells / max(stats["fill_launches"], 1)
        launch_ms = ms_fill / max(stats["fill_launches"], 1)
        achieved_gbs = per_launch_cells * PTR_BYTES_PER_CELL / (launch_ms * 1e-3) * 1e-9
        fp64_achieved = cells * FP64_INSTR_PER_CELL / 32.0 / (ms_fill * 1e-3) * 1e-9
        fp64_executed = cells * FP64_INSTR_EXECUTED / 32.0 / (ms_fill * 1e-3) * 1e-9
        line = {
            "metric": "graph_dp_gcups", "value": total_cells / (ms_dev * 1e-3) * 1e-9, "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "query placement (BASELINE configs[1]): %d reads x 150 nt per GPU vs 64-taxon x 1.5 kb "
                                   "reference alignment+tree, 1 alignment per read (fill + end corner + traceback)" % args.reads,
                       "reads_per_gpu": args.reads, "targets": info["n_targets"], "cells_per_step_per_gpu": cells,
                       "l2_policy": "inputs+outputs per step (%.1f GB of back-pointers) exceed L2" % (stats["traceback_bytes"] * 1e-9),
                       "parallelism": "independent alignments sharded by index range, %d rank(s); results + packed paths "
                                      "gathered to rank 0 over NCCL every step (%.2f ms/step on rank 0)" % (world, gather_ms / args.steps),
                       "host": "%d cores, %d packing threads per rank, rank bound to the %d cores next to its GPU" % (
                           n_cores_box, int(os.environ["PG2_PACK_THREADS"]), len(near_cores))},
            "e2e": {"value": total_cells / (ms_e2e * 1e-3) * 1e-9, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e,
                    "results_equal_resident_batch": e2e_same},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # the fill kernel is bound by the FP64 / issue port, not by HBM: `frac` is the contract figure (the
            # reference's 22 FP64-pipe instructions per cell against the chip's measured FP64 issue rate), `frac_executed`
            # the same with the instructions the kernel really issues; the HBM view of the same launch is under "hbm"
            "roofline": {"bound": "fp64_issue", "achieved": fp64_achieved, "peak": fp64_peak, "unit": "1e9 FP64-pipe warp-instr/s",
                         "frac": fp64_achieved / fp64_peak if fp64_peak else None,
                         "instr_per_cell": FP64_INSTR_PER_CELL,
                         "frac_executed": fp64_executed / fp64_peak if fp64_peak else None,
                         "instr_per_cell_executed": FP64_INSTR_EXECUTED,
                         "peak_source": "pg2_measure_fp64_issue (independent DADDs, this run; theory 148 SM x 4 x clock / 2)",
                         "candidate_update_peak": cand.value,
                         "traffic": traffic,
                         "traffic_note": "DRAM read+write bytes per fill launch, ncu capture profiles/" + TRAFFIC_PROFILE + "; "
                                         "algorithmic bytes per launch = %d" % int(per_launch_cells * PTR_BYTES_PER_CELL),
                         "kernel": "lane_fill_kernel" if stats["jobs_lanes"] >= stats["jobs_strip"] else "strip_fill_kernel",
                         "launch_ms": launch_ms,
                         "hbm": {"achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / peaks["hbm_gbs"],
                                 "peak_source": peaks_kind, "algorithmic_bytes_per_cell": PTR_BYTES_PER_CELL},
                         "issue_slots": {
                             "cycles_8dadd_16fadd_both": [mix[0], mix[1], mix[2]],
                             "busy_frac_ncu": dispatch_busy["issue_active_pct"] / 100.0 if dispatch_busy else None,
                             "fp64_pipe_frac_ncu": dispatch_busy["fp64_pipe_active_pct"] / 100.0 if dispatch_busy else None,
                             "alu_pipe_frac_ncu": dispatch_busy.get("alu_pipe_active_pct", 0.0) / 100.0 if dispatch_busy else None,
                             "warp_instr_per_32_cells_ncu": dispatch_busy["warp_instructions"] / (cells / 32.0) if dispatch_busy else None,
                             "note": "8 DADD + 16 FADD per loop iteration cost cycles[2] ~ 24 issue cycles, not cycles[0] + cycles[1]: "
                                     "FP64 instructions (2 pipe cycles each) overlap other issue, so the kernel's bound is the issue "
                                     "port (1 warp-instruction / cycle / sub-partition); busy_frac_ncu = smsp__issue_active of the "
                                     "fill launches in the committed ncu capture"}},
            "fill_ms_per_step": ms_fill, "traceback_ms_per_step": tb_ms / args.steps, "wall_ms_per_step": ms_wall,
            "jobs_ok": ok, "jobs": len(jobs),
            "kernels": {"lanes": stats["jobs_lanes"], "strip": stats["jobs_strip"], "wavefront": stats["jobs_wavefront"], "pstrip": stats["jobs_pstrip"], "band": stats["jobs_band"]},
        }
        # strong scaling: --reads alignments in TOTAL over the ranks (BASELINE configs[1] is one 100k-read job); at N=1 it is
        # the main line itself
        if strong_vals is not None:
            line["strong"] = {"reads_total": args.reads, "cells_total": strong_cells,
                              "value": strong_cells / (float(strong_vals[0]) * 1e-3) * 1e-9, "unit": "GCUPS", "ms_per_step": float(strong_vals[0]),
                              "e2e": {"value": strong_cells / (float(strong_vals[1]) * 1e-3) * 1e-9, "ms_per_step": float(strong_vals[1])}}
        elif world == 1:
            line["strong"] = {"reads_total": args.reads, "cells_total": cells, "value": line["value"], "unit": "GCUPS", "ms_per_step": ms_dev,
                              "e2e": {"value": line["e2e"]["value"], "ms_per_step": ms_e2e}}
        if not args.no_configs:
            line["configs"] = other_configs(eng)
        # the CPU leg runs at N=1 only (rank 0 would otherwise keep the other ranks waiting in a barrier) and doubles as the
        # parity check of this run
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"], line["parity_checked_jobs"] = cpu_baseline_sample(jobs, eng, results, steps_buf)
        emit(line)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
