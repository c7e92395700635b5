#!/usr/bin/env python
"""Headline benchmark: user-item pairs scored / second (fused top-k) at rank 50.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K ...  (reference CPU path, oracle port)

Workload (BASELINE.json configs[1], "C2"): synthetic 1M users x 100K items, ~0.1% nnz
(1e8 interactions, Zipf item popularity, log-normal user degrees), SVDModel rank 50,
filter_seen, top-10, every user scored against every item.  One *step* = one full
get_recommendations pass: SpMM E = P.V, fused score+mask+top-k, list merge.

Multi-GPU (weak scaling in items): every rank owns a 100K-item shard of
the item factors (total items = N x 100K); user embeddings are computed redundantly per
rank; per-shard top-k candidates are exchanged with ONE all-to-all by user range and
merged locally.  value = all (user, item) pairs of the job / max-over-ranks step time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--items", type=int, default=100_000, help="items PER GPU (weak scaling)")
    ap.add_argument("--nnz", type=int, default=100_000_000)
    ap.add_argument("--rank", type=int, default=50)
    ap.add_argument("--topk", type=int, default=10)
    ap.add_argument("--kernel", default=None, choices=[None, "simt", "tcgen05"])
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-build", action="store_true", help="random orthonormal factors instead of build()")
    return ap.parse_args()


# ------------------------------------------------------------------ data ------------
def synth_csr_torch(n_users, n_items, nnz_target, seed, device):
    """Zipf-popular items, log-normal degrees, ratings 1..5; built with torch on `device`
    (data generation is not part of any timed region).  Returns host-pinned CSR tensors."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    deg = torch.exp(torch.randn(n_users, generator=g, device=device))
    deg = torch.clamp((deg * (nnz_target / n_users / deg.mean())).round(), 1, max(1, n_items // 2)).to(torch.int64)
    w = 1.0 / torch.arange(1, n_items + 1, device=device, dtype=torch.float64)
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    perm = torch.randperm(n_items, generator=g, device=device)
    rows = torch.repeat_interleave(torch.arange(n_users, device=device), deg)
    u = torch.rand(rows.shape[0], generator=g, device=device, dtype=torch.float64)
    cols = perm[torch.searchsorted(cdf, u).clamp_(max=n_items - 1)]
    key = torch.unique(rows * n_items + cols)          # sorted by (row, col), duplicates dropped
    rows = key // n_items
    cols = (key - rows * n_items).to(torch.int32)
    counts = torch.bincount(rows, minlength=n_users)
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(counts, 0)
    vals = torch.randint(1, 6, (cols.shape[0],), generator=g, device=device).to(torch.float32)
    return indptr, cols, vals


def sample_clocks(stop, out):
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            txt = subprocess.run(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-i",
                                  os.environ.get("LOCAL_RANK", "0")], capture_output=True, text=True, timeout=5).stdout
            out.append(txt.strip().split("\n")[0])
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    sm, mx, reasons = [], [], set()
    for line in samples:
        p = [x.strip() for x in line.split(",")]
        if len(p) < 7:
            continue
        try:
            sm.append(float(p[0])); mx.append(float(p[1]))
        except ValueError:
            continue
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons)}


# ------------------------------------------------------------- CPU baseline ---------
def cpu_baseline(indptr, indices, values, n_items, v64, topk, budget_s, memory_hard_limit=1.0):
    """The oracle port of the reference's chunked driver (models.py:359-405) on the first
    chunks of users, with all host BLAS threads; returns pairs/s and what was sampled."""
    from oracle import polara_oracle as po
    n_users = len(indptr) - 1
    chunk = po.get_chunk_size((n_users, n_items), topk, 1, memory_hard_limit)
    t0 = time.perf_counter()
    done_users = 0
    n_chunks = 0
    bounds = po.range_division(n_users, chunk)
    for a, b in zip(bounds[:-1], bounds[1:]):
        lo, hi = indptr[a], indptr[b]
        user = np.repeat(np.arange(b - a), np.diff(indptr[a:b + 1]))
        item = indices[lo:hi].astype(np.int64)
        fdbk = values[lo:hi].astype(np.float64)
        # one chunk exactly as _slice_recommender does it (models.py:359-371)
        import scipy.sparse as sps
        p = sps.csr_matrix((fdbk, (user, item)), shape=(b - a, n_items))
        scores = po.svd_slice_scores(p, v64)
        po.downvote_seen_items(scores, user, item)
        po.get_topk_elements(scores, topk)
        done_users += b - a
        n_chunks += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=done_users * n_items / dt, seconds=dt, users=int(done_users), chunks=n_chunks,
                chunk_users=int(chunk))


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus
    if world > 1 and world != n_gpus:
        raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")
    n_items_total = args.items * n_gpus
    workload = "C2: synthetic %dM users x %dK items/GPU, nnz %.0e, SVD rank %d, filter_seen, top-%d" % (
        args.users // 1_000_000, args.items // 1000, args.nnz, args.rank, args.topk)
    base = {"metric": "user-item pairs scored/sec (fused top-k) at rank %d" % args.rank, "unit": "pairs/s",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": workload, "users": args.users, "items_total": n_items_total,
                       "items_per_gpu": args.items, "nnz": args.nnz, "rank": args.rank, "topk": args.topk,
                       "parallelism": "item-shard x%d" % n_gpus,
                       "l2_policy": "inputs (P, E, lists > 1 GB) larger than the 126 MB L2"}}

    if args.impl == "reference":
        if rank != 0:
            return
        run_reference(args, base, n_items_total)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from polara_b200 import _build
    _build.build()
    from polara_b200.engine import get_engine
    from polara_b200.host import ArrayData
    from polara_b200.models import B200SVDModel
    from polara_b200 import dist as pdist
    eng = get_engine(local_rank)
    if args.kernel:
        eng.set_score_kernel(args.kernel)

    # ---------------- synthetic inputs (same seed on every rank) ----------------------
    indptr_d, indices_d, values_d = synth_csr_torch(args.users, n_items_total, args.nnz, 20260924, dev)
    nnz = int(indices_d.shape[0])
    if nnz < 0.97 * args.nnz:      # duplicates of popular items were dropped: draw more to land on the target
        del indptr_d, indices_d, values_d
        indptr_d, indices_d, values_d = synth_csr_torch(args.users, n_items_total,
                                                        int(args.nnz * (args.nnz / nnz) ** 1.15), 20260924, dev)
        nnz = int(indices_d.shape[0])
    indptr_h = indptr_d.cpu().pin_memory(); indices_h = indices_d.cpu().pin_memory(); values_h = values_d.cpu().pin_memory()
    shape = (args.users, n_items_total)
    data = ArrayData(np.zeros((1, 2), dtype=np.int64), np.ones(1), shape)
    data.train_csr = (indptr_h, indices_h, values_h, shape)
    data.test_csr = ((indptr_h, indices_h, values_h), shape)      # known-user scenario: P = A
    model = B200SVDModel(data)
    model.verbose = False
    model.rank = args.rank
    model.topk = args.topk
    model.score_kernel = args.kernel
    sharder = pdist.ItemShard(rank, world, n_items_total) if world > 1 else None
    model.shard = sharder          # world > 1: row-sharded build, item-sharded scoring

    # ---------------- build() (timed once; not part of the step) ----------------------
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if args.skip_build:
        q = np.linalg.qr(np.random.default_rng(0).standard_normal((n_items_total, args.rank)))[0]
        model.factors = {"userid": None, "itemid": q, "singular_values": np.ones(args.rank)}
        model._is_ready = True
    else:
        model.build()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0

    # ---------------- device-resident step --------------------------------------------
    from polara_b200.engine import DeviceCSR
    p_dev = DeviceCSR(indptr_d, indices_d, values_d, shape)
    v_dev = model._device_factor("itemid")
    step = pdist.make_step(eng, p_dev, v_dev, args.rank, args.topk, sharder)
    for _ in range(args.warmup):
        ids = step()
    torch.cuda.synchronize()
    launches0 = eng.stats()[0]
    clocks, stop = [], threading.Event()
    th = threading.Thread(target=sample_clocks, args=(stop, clocks), daemon=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    th.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record()
    for _ in range(args.steps):
        ids = step()
        kernel_ms.append(None)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    stop.set(); th.join()
    ms = ev0.elapsed_time(ev1)
    stats = eng.stats()
    launches = stats[0] - launches0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    pairs = float(args.users) * float(n_items_total)
    value = pairs / (ms_per_step * 1e-3)

    # ---------------- dominant-kernel roofline (scoring kernel alone) ------------------
    score_ms = pdist.time_score_kernel(eng, p_dev, v_dev, args.rank, args.topk, sharder, reps=max(3, args.steps))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if peaks else "fallback 1.59 PFLOP/s"
    flops = 2.0 * args.users * (n_items_total / world) * args.rank
    achieved_tf = flops / (score_ms * 1e-3) / 1e12
    # DRAM bytes of one launch of the fused kernel from the committed ncu --set full capture (same C2 workload only)
    traffic, traffic_src = None, None
    prof = os.path.join(ROOT, "profiles", "score_topk_tc_r1_ncu.txt")
    if os.path.exists(prof) and (args.users, args.items, args.nnz, args.rank, args.topk) == (1_000_000, 100_000, 100_000_000, 50, 10) \
            and world == 1:
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        got = {}
        for line in open(prof):
            f = line.split()
            if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[1] in unit:
                got[f[0]] = float(f[2]) * unit[f[1]]
        if len(got) == 2:
            traffic, traffic_src = sum(got.values()), "profiles/score_topk_tc_r1_ncu.txt (ncu --set full, one launch of this workload)"
    roofline = {"bound": "tensor", "kernel": "fused score+mask+top-k (%s)" % (args.kernel or "default"),
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "traffic": traffic, "traffic_unit": "bytes (dram read + write)", "traffic_source": traffic_src,
                "kernel_ms": score_ms, "peak_source": peak_src,
                "algorithmic_flops_per_launch": flops}

    out = dict(base)
    out.update({"value": value, "ms_per_step": ms_per_step, "dtype": "f32 (bf16 tensor-core filter, exact fp32 rescoring)"
                if (args.kernel or "tcgen05") == "tcgen05" else "f32",
                "gpu_launches": int(launches), "roofline": roofline,
                "clocks": summarize_clocks(clocks), "build_s": build_s,
                "build_detail": model.last_timings, "nnz_actual": nnz})

    # ---------------- end to end through the model API (host buffers) -----------------
    if not args.no_e2e:
        e2e_fn = pdist.make_e2e(model, sharder)
        if os.environ.get("PB200_STREAM_CHUNKS"):
            model.stream_chunks = int(os.environ["PB200_STREAM_CHUNKS"])
        recs = None
        for _ in range(3):
            recs = e2e_fn()          # holding the previous result, like the timed loop: both pinned result blocks get cached
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 5))
        for _ in range(reps):
            recs = e2e_fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # whole-job bytes: every rank copies the row pointers, the nnz arrays cross PCIe once (sliced by rank)
        h2d = indptr_h.numel() * 8 * world + indices_h.numel() * 4 + values_h.numel() * 4
        d2h = args.users * args.topk * 8
        if os.environ.get("BENCH_DEBUG"):
            model.profile_phases = True
            for _ in range(3):
                ta = time.perf_counter()
                recs = e2e_fn()
                tb = time.perf_counter()
                recs = None
                print("e2e call %.2f ms, release %.2f ms" % ((tb - ta) * 1e3, (time.perf_counter() - tb) * 1e3), file=sys.stderr)
            print("rank", rank, "e2e phases", model.last_score_timings, file=sys.stderr)
        out["e2e"] = {"value": pairs / dt, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": int(d2h), "s_per_step": dt,
                      "call": "B200SVDModel.get_recommendations() on pinned host CSR"}

    # ---------------- CPU baseline on this box's host cores (rank 0, N=1) -------------
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        v64 = model.factors["itemid"].astype(np.float64)
        cb = cpu_baseline(indptr_h.numpy(), indices_h.numpy(), values_h.numpy(), n_items_total, v64, args.topk,
                          args.cpu_seconds)
        out["cpu_baseline"] = {"value": cb["value"], "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "first %d reference chunks (%d users of %d, chunk=%d users as "
                                         "utils.get_chunk_size gives) in %.1f s; all host BLAS threads"
                                         % (cb["chunks"], cb["users"], args.users, cb["chunk_users"], cb["seconds"])}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args, base, n_items_total):
    """Reference arm: the oracle port of the reference's CPU path on a bounded sample per step."""
    rng = np.random.default_rng(0)
    # same shape of work, generated on the host at the sample size only
    from polara_b200.synth import popularity_csr
    from oracle import polara_oracle as po
    chunk = po.get_chunk_size((args.users, n_items_total), args.topk, 1, 1.0)
    sample_users = chunk * 4
    indptr, indices, values = popularity_csr(sample_users, n_items_total, int(args.nnz * sample_users / args.users),
                                             seed=20260924)
    v64 = np.linalg.qr(rng.standard_normal((n_items_total, args.rank)))[0]

    def one_step():
        return cpu_baseline(indptr, indices, values, n_items_total, v64, args.topk, budget_s=1e9)
    for _ in range(min(args.warmup, 1)):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = one_step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sample_users * n_items_total / dt
    out = dict(base)
    out.update({"impl": "reference", "value": value, "ms_per_step": dt * 1e3, "dtype": "f64",
                "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": "%d users (4 reference chunks of %d) x %d items per step"
                                           % (sample_users, chunk, n_items_total)},
                "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
