#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on its headline config.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c5|c1] [--small]

A "step" is one BPR epoch = |R| fused sample-gather-dot-sigmoid-scatter triples in ONE kernel launch on
synthetic data of BASELINE config #2 (1M users x 100K items x 10M feedback, d=64).  For N > 1 (torchrun, one
rank per GPU) the workload is weak-scaled: every rank owns 1M users / 10M feedback, the item table is
replicated and its deltas are all-reduced (NCCL) once per epoch.

One JSON line on rank 0; see DESIGN.md "measurement" for every key.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (users per rank, items, feedback per rank, d, description)
    "c2": (1_000_000, 100_000, 10_000_000, 64, "BPR 1M users x 100K items x 10M feedback, d=64 (BASELINE configs[1])"),
    "c5": (1_250_000, 1_000_000, 25_000_000, 128, "BPR 10M x 1M x 200M, d=128 split over 8 ranks (BASELINE configs[4] per-rank share)"),
    "c1": (943, 1_682, 100_000, 16, "BPR ml-100k-shaped surrogate, d=16 (BASELINE configs[0])"),
    "small": (50_000, 10_000, 500_000, 64, "reduced smoke size (NOT a bench number)"),
}
LR, REG = 0.05, 0.01            # BPR defaults, model/cf/model.go:391-392
INIT_STD = 0.001                # :394
E2E_EPOCHS = 100                # BPR default NEpochs, model/cf/model.go:390
ZIPF_S = 1.0                    # item popularity skew of the synthetic feedback
_CPU_CACHE = {}


def bytes_per_triple(d):
    return 6 * d * 4 + 12       # 3 rows read + 3 rows written + 3 ids


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return None
        sm = sorted(float(r[0]) for r in rows)
        reasons = [n for k, n in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap"))
                   if any(r[k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


_DATA_CACHE = {}


def get_data(wl, seed):
    from gorse_b200 import synth

    if (wl, seed) not in _DATA_CACHE:
        upr, n_items, fpr, d, _ = WORKLOADS[wl]
        _DATA_CACHE[(wl, seed)] = synth.make_feedback(upr, n_items, fpr, seed=seed, zipf_s=ZIPF_S, exact=True)
    return _DATA_CACHE[(wl, seed)]


def make_shard(wl, rank, world):
    """This rank's rows of the global user CSR (the C ABI takes ONLY the rank's own rows in a distributed context):
    users [rank*upr, (rank+1)*upr) of n_users = upr*world, seeded per rank."""
    upr, n_items, fpr, d, _ = WORKLOADS[wl]
    off_l, items = get_data(wl, 1000 + rank)
    return upr * world, n_items, d, off_l, items


def cpu_baseline(wl, seconds_budget=3.0, prefer_ref=True):
    """The reference's CPU path for the same step on this box's host cores: per triple the reference's own call
    sequence (2x dot, exp, 3 row copies, 10 vector ops, model/cf/model.go:469-488) through the reference's C kernels
    (oracle/_ref, compiled from /root/reference/common/floats/src) when they are present and the CPU has AVX-512,
    else through the oracle port; Hogwild over all host cores.  Bounded sample of the workload's step."""
    import oracle
    from gorse_b200 import synth

    upr, n_items, fpr, d, _ = WORKLOADS[wl]
    if wl not in _CPU_CACHE:
        off, items = get_data(wl, 1000)
        rng = np.random.default_rng(0)
        P = (rng.standard_normal((upr, d)) * INIT_STD).astype(np.float32)
        Q = (rng.standard_normal((n_items, d)) * INIT_STD).astype(np.float32)
        _CPU_CACHE[wl] = (off, items, P, Q, np.nonzero(np.diff(off) > 0)[0].astype(np.int32))
    off, items, P, Q, active = _CPU_CACHE[wl]
    cores = len(os.sched_getaffinity(0))
    kind = "port"
    use_ref = False
    if prefer_ref and oracle.ref_available() and "avx512f" in open("/proc/cpuinfo").read() and oracle.ref_bind():
        kind, use_ref = "reference", True
    probe = 200_000
    sec = oracle.bpr_epoch_threads(P, Q, off, items, active, 1, probe, LR, REG, cores, use_ref)
    n = int(min(off[-1] * 4, max(probe, probe * seconds_budget / max(sec, 1e-6))))
    sec = oracle.bpr_epoch_threads(P, Q, off, items, active, 2, n, LR, REG, cores, use_ref)
    return {"value": n / sec, "unit": "triples/s", "cores": cores, "kind": kind,
            "sample": f"{n} sampled triples of the {WORKLOADS[wl][4]} step, Hogwild over {cores} threads, {sec:.2f} s wall"}, n, sec


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    wl = args.workload
    upr, n_items, fpr, d, desc = WORKLOADS[wl]
    times, vals, last = [], [], None
    for s in range(args.warmup + args.steps):
        cb, n, sec = cpu_baseline(wl, seconds_budget=2.0)
        if s >= args.warmup:
            times.append(sec)
            vals.append(cb["value"])
        last = cb
    v = float(np.mean(vals))
    last["value"] = v
    line = {"impl": "reference", "metric": "BPR-MF triples/sec", "value": v, "unit": "triples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "users": upr, "items": n_items, "feedback": fpr, "d": d, "lr": LR, "reg": REG,
                       "step": "bounded sample of one epoch on host cores"},
            "cpu_baseline": last,
            "e2e": {"value": v, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    rank, world, local = dist_env()
    assert world == args.gpus or world == 1, f"WORLD_SIZE {world} != --gpus {args.gpus}"
    dist = None
    if world > 1:
        import torch  # control plane only: rendezvous, id broadcast, max-over-ranks
        import torch.distributed as dist

        dist.init_process_group("gloo", rank=rank, world_size=world)
    import gorse_b200 as gb

    wl = args.workload
    upr, n_items_w, fpr, d, desc = WORKLOADS[wl]
    n_users, n_items, d, off, items = make_shard(wl, rank, world)
    n_local = int(off[-1])
    if world > 1:
        import torch

        cnt = torch.tensor([n_local], dtype=torch.int64)
        allc = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allc, cnt)
        # every rank runs the same number of steps per epoch (the global step stream is split evenly)
        steps_per_epoch = int(min(int(c.item()) for c in allc)) * world
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idbuf = torch.frombuffer(bytearray(gb.nccl_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(idbuf, 0)
        ctx = gb.Context(local, rank, world, bytes(idbuf.numpy().tobytes()))
    else:
        steps_per_epoch = n_local
        ctx = gb.Context(local)

    def max_over_ranks(x):
        if world == 1:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    model = gb.CFModel(ctx, n_users, n_items, d, off, items)
    model.init_normal(0.0, INIT_STD, 0)
    ctx.sync()
    scatter = gb.SCATTER_ATOMIC if args.scatter == "atomic" else gb.SCATTER_STORE

    # ---- device-resident timing: K epochs, CUDA events on the library's own stream, max over ranks ----
    for w in range(args.warmup):
        model.bpr_epoch(LR, REG, steps_per_epoch, 100 + w, scatter)
    ctx.barrier()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launch_count()
    t0 = time.time()
    ctx.timer_begin()
    for s in range(args.steps):
        model.bpr_epoch(LR, REG, steps_per_epoch, 1000 + s, scatter)
    ctx.barrier() if world > 1 else None
    ms_total = ctx.timer_end()
    t1 = time.time()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_total = max_over_ranks(ms_total)
    ms_per_step = ms_total / args.steps
    value = steps_per_epoch / (ms_per_step * 1e-3)

    # ---- the dominant kernel alone (one epoch launch per event pair), for the roofline ----
    kms = []
    for s in range(min(args.steps, 5)):
        ctx.timer_begin()
        model.bpr_epoch(LR, REG, steps_per_epoch, 2000 + s, scatter)  # N > 1: includes the delta exchange
        kms.append(ctx.timer_end())
    k_ms = float(np.mean(kms))
    peak, peak_src = read_peaks()
    per_launch_triples = steps_per_epoch / world
    achieved = bytes_per_triple(d) * per_launch_triples / (k_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "bpr_epoch_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"{wl}_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "kernel": f"one epoch = bpr_epoch_kernel<{d // 16 if d % 16 == 0 else 0},{args.scatter}> (free-running, ~47 % of the step at "
                          "Zipf 1.0) + bpr_hot_apply_kernel (~50 %) + 5 tiny queue kernels" + (" + Q delta all-reduce" if world > 1 else ""),
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": bytes_per_triple(d) * per_launch_triples,
                "peak_source": peak_src}

    # ---- end to end through the C-ABI with HOST buffers: what cf.BPR.Fit does per call ----
    # create (CSR H2D) + factor upload from the pinned mirror + E2E_EPOCHS epochs + factor download, all timed
    e2e, e2e_error = None, None
    if not args.no_e2e:
        hp = hq = None
        try:
            # page-locked host mirror: this rank's user rows + the item table (the shim's flat pinned mirror)
            hp = gb.PinnedArray((upr, d))
            hq = gb.PinnedArray((n_items, d))
        except Exception as ex:
            e2e_error = f"{type(ex).__name__}: {ex}"
        # every rank must take the same path (bpr_epoch is collective for N > 1)
        all_ok = -max_over_ranks(-1.0 if (hp is not None and hq is not None) else 0.0) > 0.5
        if not all_ok:
            e2e_error = e2e_error or "another rank could not allocate its pinned mirror"
            for buf in (hp, hq):
                if buf is not None:
                    buf.free()
    if not args.no_e2e and e2e_error is None:
        m2 = None
        try:
            rng = np.random.default_rng(1)
            hp.array[:] = (rng.standard_normal((upr, d)) * INIT_STD).astype(np.float32)
            hq.array[:] = (np.random.default_rng(2).standard_normal((n_items, d)) * INIT_STD).astype(np.float32)
            # the C ABI takes the base of the FULL user table and touches only this rank's rows [rank*upr, (rank+1)*upr)
            p_base = C.c_void_p(hp.array.ctypes.data - rank * upr * d * 4)
            q_ptr = gb.ptr(hq.array)
            e_epochs = max(1, args.e2e_epochs)
            model.close()
            model = None
            if dist:
                dist.barrier()
            w0 = time.time()
            m2 = gb.CFModel(ctx, n_users, n_items, d, off, items)
            gb.check(gb.lib.gorse_b200_cf_set_factors(m2.h, p_base, q_ptr))
            for s in range(e_epochs):
                m2.bpr_epoch(LR, REG, steps_per_epoch, 3000 + s, scatter)
            gb.check(gb.lib.gorse_b200_cf_get_factors(m2.h, p_base, q_ptr))
            w1 = time.time()
            e_sec = max_over_ranks(w1 - w0)
            # the verdict is taken collectively so that every rank leaves this block the same way
            finite = bool(np.isfinite(hq.array).all() and np.isfinite(hp.array).all())
            if max_over_ranks(0.0 if finite else 1.0) > 0.5:
                raise FloatingPointError("non-finite factors after the end-to-end leg")
            h2d = (off.nbytes + items.nbytes + upr * d * 4 + n_items * d * 4)
            d2h = upr * d * 4 + n_items * d * 4
            e2e = {"value": e_epochs * steps_per_epoch / e_sec, "unit": "triples/s",
                   "h2d_bytes_per_step": h2d / e_epochs, "d2h_bytes_per_step": d2h / e_epochs,
                   "what": f"one Fit-shaped call per rank: cf_create (CSR upload) + set_factors from the pinned mirror + {e_epochs} epochs "
                           f"(reference default NEpochs) + get_factors; bytes are per rank and per epoch, amortised over the call",
                   "wall_s": e_sec}
        except Exception as ex:  # keep the contract line even if the end-to-end leg fails
            e2e_error = f"{type(ex).__name__}: {ex}"
        finally:
            if m2 is not None:
                m2.close()
            hp.free()
            hq.free()
    if model is not None:
        model.close()

    line = None
    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu:
            cb, _, _ = cpu_baseline(wl)
        line = {"metric": "BPR-MF triples/sec", "value": value, "unit": "triples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc + (f" per rank x {world} ranks (weak scaling; Q replicated, per-epoch NCCL all-reduce of item deltas)" if world > 1 else ""),
                           "users": n_users, "items": n_items, "feedback_per_epoch": steps_per_epoch, "d": d, "lr": LR, "reg": REG,
                           "item_popularity": f"zipf({ZIPF_S})", "scatter": args.scatter,
                           "cache": "user table (256 MB/rank at c2) is larger than L2 (126 MB); the item table is meant to stay L2-resident; no flush between steps"},
                "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        if e2e_error:
            line["e2e_error"] = e2e_error
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def run_als(args):
    """BASELINE configs[2]: eALS ("CCD") on the C2 data at d = 128; one step = one epoch (both half-sweeps + 2 Grams)."""
    import gorse_b200 as gb
    import oracle

    U, I, R, d = (1_000_000, 100_000, 10_000_000, 128) if not args.small else (50_000, 10_000, 500_000, 128)
    from gorse_b200 import synth
    off, items = synth.make_feedback(U, I, R, seed=1000, zipf_s=ZIPF_S, exact=True)
    ioff, iusers = gb.transpose_csr(off, items, I)
    reg, alpha = 0.06, 0.001  # ALS defaults, model/cf/model.go:584-585
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, d, off, items, ioff, iusers) as m:
        m.init_normal(0.0, 0.1, 0)
        for _ in range(args.warmup):
            m.als_epoch(reg, alpha)
        sampler = ClockSampler(0)
        l0, t0 = ctx.launch_count(), time.time()
        ctx.timer_begin()
        for _ in range(args.steps):
            m.als_epoch(reg, alpha)
        ms = ctx.timer_end() / args.steps
        t1 = time.time()
        launches = ctx.launch_count() - l0
        clocks = sampler.stop(t0, t1)
    peak, src = read_peaks()
    nbytes = 2 * R * (4 * d + 4) + 3 * (U + I) * 4 * d
    cb = None
    if not args.no_cpu:
        cores = len(os.sched_getaffinity(0))
        Us, Is, Rs = 50_000, 10_000, 500_000
        o2, i2 = synth.make_feedback(Us, Is, Rs, seed=1000, zipf_s=ZIPF_S, exact=True)
        io2, iu2 = gb.transpose_csr(o2, i2, Is)
        rng = np.random.default_rng(0)
        P = (rng.standard_normal((Us, d)) * 0.1).astype(np.float32)
        Q = (rng.standard_normal((Is, d)) * 0.1).astype(np.float32)
        sec = oracle.als_epoch_threads(P, Q, o2, i2, io2, iu2, reg, alpha, cores)
        cb = {"value": 2 * Rs / sec, "unit": "feedback visits/s", "cores": cores, "kind": "port",
              "sample": f"one epoch at {Us} x {Is} x {Rs}, d={d}: rows over {cores} threads, serial Gram as in the reference, {sec:.2f} s"}
    print(json.dumps({"metric": "eALS (CCD) feedback visits/sec", "value": 2 * R / (ms * 1e-3), "unit": "feedback visits/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": f"eALS {U} users x {I} items x {R} feedback, d={d} (BASELINE configs[2])", "reg": reg, "alpha": alpha,
                                 "item_popularity": f"zipf({ZIPF_S})"},
                      "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": nbytes / (ms * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes_per_epoch": nbytes,
                                   "peak_source": src},
                      "cpu_baseline": cb, "e2e": None, "gpu_launches": int(launches), "clocks": clocks}), flush=True)


def run_topk(args):
    """BASELINE configs[3]: all-pairs top-100 over 1M x 128 unit vectors (cosine via -dot); value = query vectors/s."""
    import gorse_b200 as gb
    import oracle

    N, d, k = (1_000_000, 128, 100) if not args.small else (100_000, 128, 100)
    nq = min(N, args.queries)
    # N > 1: the vectors are replicated, rank r answers its own nq query rows, no collective (weak scaling)
    rank, world, local = dist_env()
    assert world == args.gpus or world == 1, f"WORLD_SIZE {world} != --gpus {args.gpus}"
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        dist.init_process_group("gloo", rank=rank, world_size=world)
    q_lo = (rank * nq) % max(1, N - nq + 1)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, d), dtype=np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    with gb.Context(local) as ctx, gb.BruteforceIndex(ctx, d, gb.METRIC_NEG_DOT) as ix:
        ix.add(X)
        ix.search_range(0, 512, k)  # builds the bf16 mirror
        # results land in page-locked host buffers, like the shim's pinned mirror (gorse_b200_host_alloc)
        pin = (gb.PinnedArray((nq, k), np.int32), gb.PinnedArray((nq, k), np.float32), gb.PinnedArray((nq,), np.int32))
        out = (pin[0].array, pin[1].array, pin[2].array)
        for _ in range(max(0, args.warmup - 1)):
            ix.search_range(q_lo, q_lo + nq, k, out=out)
        if dist:
            dist.barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        ix.debug_stage1()
        l0, t0 = ctx.launch_count(), time.time()
        ms_list = []
        for s in range(args.steps):
            ctx.timer_begin()
            idx, dist_, cnt = ix.search_range(q_lo, q_lo + nq, k, out=out)   # host buffers out: this IS the end-to-end call
            ms_list.append(ctx.timer_end())
        t1 = time.time()
        launches = ctx.launch_count() - l0
        fb = ix.debug_fallback_rows()
        s1_ms, s1_flop = ix.debug_stage1()
        clocks = sampler.stop(t0, t1) if sampler else None
    ms = float(np.mean(ms_list))
    assert cnt.min() == k
    cnt = None
    for p_ in pin:
        p_.free()
    if dist:
        import torch

        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # device-timed, max over ranks
        ms = float(t.item())
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
    nq_all = nq * world
    flop = 2.0 * nq_all * N * d
    peak = 1427.2
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak = float(json.load(open(pp)).get("bf16_tflops_sustained", peak))
    cb = None
    if not args.no_cpu and world == 1:
        cores = len(os.sched_getaffinity(0))
        sample = 4 * cores
        _, _, _, sec = oracle.bruteforce_all(X, 0, sample, k, metric=oracle.METRIC_NEG_DOT, n_threads=cores)
        cb = {"value": sample / sec, "unit": "vectors/s", "cores": cores, "kind": "port",
              "sample": f"{sample} queries of the same 1M set: reference-order dot + Go heap per query, {cores} threads, {sec:.2f} s"}
    print(json.dumps({"metric": "item-to-item top-k vectors/sec", "value": nq_all / (ms * 1e-3), "unit": "vectors/s", "n_gpus": world, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "bf16 tensor-core candidate generation + f32 exact re-rank", "data": "synthetic",
                      "config": {"workload": f"all-pairs top-{k} over {N} x {d} unit vectors, {nq} query rows per step and rank (BASELINE configs[3])"
                                             + (f" x {world} ranks, vectors replicated, queries sharded, no collective" if world > 1 else ""),
                                 "metric": "-dot (cosine on unit vectors)", "fallback_rows": int(fb)},
                      "roofline": {"bound": "tensor", "achieved": s1_flop / (s1_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                                   "frac": s1_flop / (s1_ms * 1e-3) / 1e12 / peak, "traffic": None, "kernel": "mma::topk_mma_kernel<4>",
                                   "kernel_ms": s1_ms / args.steps, "algorithmic_flop_per_launch": s1_flop / args.steps,
                                   "whole_call_tflops": flop / (ms * 1e-3) / 1e12,
                                   "note": "dominant kernel = the tcgen05 sweep (CUDA events on the library's stream); algorithmic flop 2*nq*N*d; "
                                           "peak = measured sustained bf16 (MEASURED_PEAKS.json). whole_call adds query mirror, prune, exact re-rank, fallback and result D2H"},
                      "cpu_baseline": cb, "e2e": {"value": nq_all / (ms * 1e-3), "unit": "vectors/s", "h2d_bytes_per_step": 0,
                                                  "d2h_bytes_per_step": nq * k * 8 + nq * 4},
                      "gpu_launches": int(launches), "clocks": clocks}), flush=True)


def build_parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c3", "c4"])
    ap.add_argument("--small", action="store_true", help="c3/c4 at reduced size (not a bench number)")
    ap.add_argument("--queries", type=int, default=151552, help="c4: query rows per step")
    ap.add_argument("--scatter", default="atomic", choices=["atomic", "store"])
    ap.add_argument("--e2e-epochs", type=int, default=E2E_EPOCHS)
    ap.add_argument("--zipf", type=float, default=None, help="item popularity exponent of the synthetic data (default 1.0; 0 = uniform)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap


def main():
    args = build_parser().parse_args()
    if args.zipf is not None:
        global ZIPF_S
        ZIPF_S = args.zipf
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "c3":
        run_als(args)
    elif args.workload == "c4":
        run_topk(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
