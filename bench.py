#!/usr/bin/env python
"""Headline benchmark: train samples/sec (fwd + bwd + update) of the embedding hot path on B200.

Workload (BASELINE.json configs[1], the config the metric is quoted on): FFM + FM second-order on synthetic
Criteo-shaped data -- 26 sparse fields, D = 16, batch 65536 per GPU, the public Criteo-Kaggle cardinalities
(33.76 M rows; FFM table 56.2 GB, FM table 2.2 GB), uniform ids, Bernoulli(0.3) labels, plain SGD.
One "step" = one FM train step + one FFM train step over the same batch through the public API
(model -> BCELoss -> backward -> FusedRowOptimizer.step(), driven by Trainer.train_loop).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--dist uniform|zipf] [--light]

Prints ONE JSON line (contract in the task brief): value = whole-job samples/s with inputs resident in HBM;
e2e = same metric with the ids/labels copied from pinned host memory and the loss read back every step;
roofline = algorithmic bytes of the dominant kernel / its CUDA-event duration vs MEASURED_PEAKS.json;
cpu_baseline = the oracle port timed on this box's host cores on a bounded sample.
`--impl reference` times the reference's CPU implementation of the path (the oracle port: torch CPU fp32,
nn.Embedding-style dense gradient + dense SGD) on the host cores.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CRITEO = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992, 5461306, 10, 5652,
          2173, 4, 7046547, 18, 15, 286181, 105, 142572]
F, D, BATCH = 26, 16, 65536
LR = 0.05
WORKLOAD = "C2: FM second-order + FFM train step, 26 Criteo-shaped sparse fields, D=16, batch 65536, SGD"
# dram__bytes_read.sum + dram__bytes_write.sum of one seg_stream_kernel launch on the full uniform-id config
# (ncu --set full, profiles/r1_ncu_full_c2.md): 2.84 GB Jacobian stash + the ~0.53 M unique table rows read and written
NCU_TRAFFIC_SEG_STREAM = 4.616e9
# algorithmic bytes per sample (fp32 rows, int64 ids, no duplicate reuse)
BYTES = {
    "fm_fwd": 26 * 64 + 208 + 4, "fm_bwd_upd": 1664 + 2 * 1664 + 208,
    "ffm_fwd": 26 * 26 * 64 + 208, "ffm_bwd_upd": 43264 + 2 * 43264,
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def make_ids(cards, B, seed, dist, device):
    g = torch.Generator(device=device).manual_seed(seed)
    cols = []
    for c in cards:
        if dist == "zipf":  # inverse-CDF of a continuous power law with exponent 1.05, truncated to [1, c]
            u = torch.rand(B, generator=g, device=device, dtype=torch.float64)
            a = 1.05
            x = ((c ** (1 - a) - 1) * u + 1) ** (1 / (1 - a))
            cols.append((x.floor().long() - 1).clamp_(0, c - 1))
        else:
            cols.append(torch.randint(0, c, (B,), generator=g, device=device))
    ids = torch.stack(cols, dim=1).contiguous()
    y = (torch.rand(B, 1, generator=g, device=device) < 0.3).float()
    return ids, y


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML every few ms on a background thread while the timed
    region runs (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, period_s=float(os.environ.get("RS_BENCH_CLOCK_PERIOD", "0.004"))):
        import threading
        self.sm, self.mask, self.max_mhz, self.ok = [], 0, None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            return
        self.period = period_s
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        while not self._stop.is_set():
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if not self.ok:
            return out
        self._stop.set()
        self.t.join(timeout=2)
        if self.sm:
            v = sorted(self.sm)
            out.update(sm_mhz=v[len(v) // 2], samples=len(v), reasons=sorted(n for bit, n in self.REASONS.items() if self.mask & bit))
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_step_factory(B, cap, threads):
    """Oracle port of the C2 step on a bounded sample: B samples, cardinalities capped at `cap` rows per field."""
    from oracle import nfield as onf
    torch.set_num_threads(threads)
    cards = [min(c, cap) for c in CRITEO]
    offsets = torch.tensor([sum(cards[:i]) for i in range(F)])
    g = torch.Generator().manual_seed(0)
    total = sum(cards)
    fm_t = torch.randn(total, D, generator=g) * 0.01
    ffm_t = torch.randn(total, F * D, generator=g) * 0.01
    fm_b, ffm_b = torch.zeros(1), torch.zeros(1)
    ids = torch.stack([torch.randint(0, c, (B,), generator=g) for c in cards], dim=1)
    y = (torch.rand(B, generator=g) < 0.3).float()

    def step():
        onf.train_step("fm", fm_t, fm_b, ids, offsets, y, LR)
        _, loss = onf.train_step("ffm", ffm_t, ffm_b, ids, offsets, y, LR, fast=True)
        return float(loss)

    return step, f"B={B} of {BATCH}, cardinalities capped at {cap} rows/field ({total} rows), dense-gradient SGD as nn.Embedding+optim.SGD"


def time_cpu(steps, warmup, B=4096, cap=20000):
    threads = os.cpu_count() or 1
    step, sample = cpu_step_factory(B, cap, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return B / dt, dt * 1e3, threads, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, ms, threads, sample = time_cpu(max(1, args.steps), max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": "train samples/sec (fwd+bwd+update)", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "impl": "CPU oracle port (torch CPU fp32, all host threads)"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch.distributed as dist
    from deeplearningrecommendationsystem_b200 import ops
    from deeplearningrecommendationsystem_b200.nfield import FieldFFM, FieldFM
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cards = [min(c, 1 << 17) for c in CRITEO] if args.light else CRITEO
    B = args.batch

    # N > 1: tables are row-sharded over the ranks (global row r on rank r % N); every step exchanges the
    # deduplicated ids / rows / row-gradients with three all-to-alls (dist.RowExchange)
    fm = FieldFM(cards, D, fused=True, seed=1, device=dev, sharded=world > 1)
    ffm = FieldFFM(cards, D, fused=True, seed=2, device=dev, sharded=world > 1)
    loss_fn = torch.nn.BCELoss()
    trainers = []
    for m in (ffm, fm):        # FFM first: the batch's sort (shared by both models) then overlaps the long FFM forward
        opt = FusedRowOptimizer(m, torch.optim.SGD([m.bias], lr=LR), lr=LR, kind="sgd")
        trainers.append(Trainer(m, loss_fn, opt))

    # a pool of distinct batches so no step re-reads the previous step's rows from L2.  8 batches > the 4 dedup memo
    # slots (LRU), so every step sorts a batch it has not seen recently: only the FM/FFM sharing inside a step hits.
    pool = [make_ids(cards, B, 1234 + rank * 100 + k, args.dist, dev) for k in range(8)]
    host_pool = [(i.cpu().pin_memory(), y.cpu().pin_memory()) for i, y in pool]

    def step(ids, y):
        for tr in trainers:
            tr.train_loop(ids, train_rating=y)
        return trainers[0].train_loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:                      # NCCL sets its peer connections up lazily: prime them outside the W warm-up steps
        for k in range(3):
            step(*pool[k % len(pool)])
    for k in range(args.warmup):
        step(*pool[k % len(pool)])
    ops.check_status(dev)

    # ---- timed: device-resident inputs
    ops.PROFILE = []
    clocks = ClockSampler(local)
    sync()
    launches0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        step(*pool[k % len(pool)])
    e1.record()
    sync()
    ms_total = e0.elapsed_time(e1)
    gpu_launches = ops.launches() - launches0
    prof, ops.PROFILE = ops.PROFILE, None

    # The end-to-end leg replays the same step as a CUDA graph (graph.GraphedTrainStep, single GPU): identical kernels,
    # but two launches of host work per step, so the loop cannot become host bound on a noisy box.  The device-resident
    # leg above stays eager so that its kernels can be bracketed with CUDA events and counted.
    graphed = None
    if args.graph and world == 1:
        from deeplearningrecommendationsystem_b200.graph import GraphedTrainStep
        for tr in trainers:     # drop the eager leg's autograd graphs: their AccumulateGrad nodes are bound to its stream
            tr.predictions_train = tr.train_loss = None
        graphed = [GraphedTrainStep(tr.model, loss_fn, tr.optimizer, warmup=1) for tr in trainers]
        eager_step = step

        def step(ids, y):                               # noqa: F811
            loss = None
            for gs in graphed:
                _, loss = gs(ids, rating=y)
            return loss
        for k in range(3):                              # untimed: one eager call, the capture, one replay
            step(*pool[k % len(pool)])
        del eager_step

    # ---- timed: end to end from pinned host memory, loss read back every step.  The next batch's ids/labels are
    # copied on a side stream while the current step computes (every copy is still inside the timed region).
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    # two device staging slots, reused (no allocation inside the timed loop)
    slots = [(torch.empty_like(pool[0][0]), torch.empty_like(pool[0][1])) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    h2d_events = []

    def fetch(k):
        hi, hy = host_pool[k % len(host_pool)]
        b = k & 1
        with torch.cuda.stream(copy_stream):
            if k >= 2:
                copy_stream.wait_event(consumed[b])        # the step that read this slot has finished
            c0 = torch.cuda.Event(enable_timing=True)
            c0.record(copy_stream)
            slots[b][0].copy_(hi, non_blocking=True)
            slots[b][1].copy_(hy, non_blocking=True)
            c1 = torch.cuda.Event(enable_timing=True)
            c1.record(copy_stream)
            copied[b].record(copy_stream)
            h2d_events.append((c0, c1))

    # every step's loss is copied back into pinned host memory (non-blocking, one slot per step) and read after the
    # final synchronise, so the host keeps enqueueing work instead of stalling on a 4-byte read each step
    loss_host = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
    sync()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    fetch(0)
    for k in range(args.steps):
        b = k & 1
        if k + 1 < args.steps:
            fetch(k + 1)
        main.wait_event(copied[b])
        loss_host[k:k + 1].copy_(step(*slots[b]).detach().reshape(1), non_blocking=True)
        consumed[b].record(main)
    t1.record()
    sync()
    loss_val = float(loss_host[-1])
    h2d_ms = sorted(a.elapsed_time(b) for a, b in h2d_events)
    ms_e2e = t0.elapsed_time(t1)
    clk = clocks.stop()

    times = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = times.tolist()
    value = world * B * args.steps / (ms_total / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        # per-kernel durations recorded around the C-ABI calls (CUDA events on the launching stream)
        agg = {}
        for name, a, b in prof:
            agg.setdefault(name, []).append(a.elapsed_time(b))
        kern = {k: sum(v) / len(v) for k, v in agg.items()}
        peak, which = peaks()
        dom = "segment_update[w416]"
        roof = None
        if dom in kern:
            achieved = BYTES["ffm_bwd_upd"] * B / (kern[dom] / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": "rs_segment_update on FFM rows (416 floats): seg_stream_kernel<SGD> + scale gather + combine", "achieved": achieved,
                    "peak": peak, "peak_source": which, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": NCU_TRAFFIC_SEG_STREAM if (not args.light and args.dist == "uniform" and B == BATCH) else None,
                    "algorithmic_bytes_per_launch": BYTES["ffm_bwd_upd"] * B, "ms_per_launch": kern[dom],
                    "note": "algorithmic bytes charge a table-row read+write per LOOKUP (no duplicate reuse, SURVEY 8d); "
                            "DRAM traffic is per UNIQUE row, hence frac > 1 while traffic/time stays below the peak"}
        extra = {}
        for name, key in (("ffm_fwd", "ffm_fwd"), ("fields_fwd", "fm_fwd"), ("segment_update[w16]", "fm_bwd_upd")):
            if name in kern:
                extra[name] = {"ms": kern[name], "algorithmic_GBps": BYTES[key] * B / (kern[name] / 1e3) / 1e9}
        for name in kern:
            if name not in extra and name != dom:
                extra[name] = {"ms": kern[name], "calls_per_step": len(agg[name]) / args.steps}
        line = {
            "metric": "train samples/sec (fwd+bwd+update)", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "batch_per_gpu": B, "fields": F, "dim": D, "rows": sum(cards), "ids": args.dist, "light": bool(args.light),
                       "l2": "8 distinct id batches cycled; per-step traffic (>8 GB) far exceeds the 126 MB L2",
                       "step_api": "trainer.Trainer.train_loop (eager)",
                       "parallelism": f"dp{world}" + ("" if world == 1 else ": batch split, tables row-sharded, dedup all-to-all of ids/rows/grads")},
            "roofline": roof, "kernels": extra,
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": B * F * 8 + B * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "last_loss": loss_val,
                    "step_api": "graph.GraphedTrainStep (CUDA graph replay)" if graphed is not None else "trainer.Trainer.train_loop (eager)",
                    "h2d_copy_ms": {"median": h2d_ms[len(h2d_ms) // 2], "max": h2d_ms[-1]}},
            "gpu_launches": gpu_launches, "clocks": clk,
        }
        if world == 1 and not args.no_cpu:
            val, ms, threads, sample = time_cpu(2, 1)
            line["cpu_baseline"] = {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ other configs
def run_other(args):
    """Non-headline BASELINE.json configs (one JSON line per model): c3 PNN-inner + AFM (F=39, D=32, B=32768),
    c4 DIN + DIEN (L=100, D=64, B=8192, 1 M items), c5 MF with two 100 M-row tables (row-sharded when N > 1)."""
    import torch.distributed as dist
    from deeplearningrecommendationsystem_b200 import ops
    from deeplearningrecommendationsystem_b200.model import DIEN, DIN
    from deeplearningrecommendationsystem_b200.nfield import FieldAFM, FieldMF, FieldPNN
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    loss_fn = torch.nn.BCELoss()
    jobs = []
    if args.workload == "c3":
        cards, B = CRITEO + [64] * 13, 32768

        def batches(k):
            return [make_ids(cards, B, 900 + rank * 100 + i, args.dist, dev) for i in range(k)]
        for name, m in (("PNN-inner", FieldPNN(cards, 32, [256, 128, 64, 32], seed=1, device=dev, sharded=world > 1)),
                        ("AFM", FieldAFM(cards, 32, 64, seed=2, device=dev, sharded=world > 1))):
            dense = [p for p in m.parameters() if p.requires_grad]
            opt = FusedRowOptimizer(m, torch.optim.SGD(dense, lr=LR), lr=LR)
            pool = batches(4)
            jobs.append((name, m, opt, [((i,), y) for i, y in pool], B, "F=39, D=32, fused sparse SGD on 33.8 M rows"))
    elif args.workload == "c4":
        B, L, items = 8192, 100, 1_000_000
        for name, cls in (("DIN", DIN), ("DIEN", DIEN)):
            m = cls(items, 64).to(dev)
            opt = torch.optim.SGD(m.parameters(), lr=0.01)
            pool = []
            for _ in range(4):
                u = torch.rand(B, L, generator=g, device=dev, dtype=torch.float64)
                hist = ((((items ** (1 - 1.05) - 1) * u + 1) ** (1 / (1 - 1.05))).floor().long() - 1).clamp_(0, items - 1)
                tgt = torch.randint(0, items, (B,), generator=g, device=dev)
                pool.append(((hist, tgt), (torch.rand(B, 1, generator=g, device=dev) < 0.3).float()))
            jobs.append((name, m, opt, pool, B, "L=100, D=64, 1 M-row item table, dense-gradient SGD (reference semantics)"))
    else:
        rows, B = (100_000_000 if not args.light else 1_000_000), 65536
        m = FieldMF(rows, rows, 64, seed=3, device=dev, sharded=world > 1)
        opt = FusedRowOptimizer(m, None, lr=LR)
        pool = []
        for _ in range(8):
            u = torch.randint(0, rows, (B,), generator=g, device=dev)
            i = torch.randint(0, rows, (B,), generator=g, device=dev)
            pool.append(((u, i), (torch.rand(B, generator=g, device=dev) < 0.3).float()))
        jobs.append(("MF", m, opt, pool, B, f"2 x {rows} rows, D=64, fused sparse SGD" + (", row-sharded all-to-all" if world > 1 else "")))

    for name, m, opt, pool, B, note in jobs:
        tr = Trainer(m, loss_fn, opt)
        if args.graph and world == 1:
            from deeplearningrecommendationsystem_b200.graph import GraphedTrainStep
            gs = GraphedTrainStep(m, loss_fn, opt)

            class _G:                                   # same surface as Trainer for the loop below
                train_loss = None

                def train_loop(self, *a, train_rating):
                    _, self.train_loss = gs(*a, rating=train_rating)
            tr = _G()
            note += ", CUDA-graph replay"
        for k in range(max(args.warmup, 4) + (5 if world > 1 else 0)):
            tr.train_loop(*pool[k % len(pool)][0], train_rating=pool[k % len(pool)][1])
        ops.check_status(dev)
        ops.PROFILE = []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            tr.train_loop(*pool[k % len(pool)][0], train_rating=pool[k % len(pool)][1])
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = ms.item() / args.steps
        if rank == 0:
            agg = {}
            for kn, a, b in prof:
                agg.setdefault(kn, []).append(a.elapsed_time(b))
            print(json.dumps({"metric": "train samples/sec (fwd+bwd+update)", "value": world * B / (ms / 1e3), "unit": "samples/s",
                              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                              "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": f"{args.workload}: {name}", "batch_per_gpu": B, "note": note, "ids": args.dist},
                              "kernels": {kn: {"ms": sum(v) / len(v), "calls_per_step": len(v) / args.steps} for kn, v in agg.items()},
                              "last_loss": tr.train_loss.item()}))
        del tr, m, opt, pool
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


def run_rank(args):
    """SURVEY.md 8(f).1 -- catalogue ranking (`recommendation()`), one JSON line per case.  Not the headline metric.

    mf:      fused rs_mf_rank over a (users x items) catalogue vs cuBLAS matmul + torch.topk (library route) and the
             numpy oracle on the host (bounded sample of users).
    deepfm:  DeepFM.recommendation(943, frame, 1682) -- the scripts' call (scripts/deepfm.py:67) -- wall clock including
             the host-side grouping of the frame, vs the reference's one-forward-per-user loop on a sample of users."""
    import time
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import catalogue_frame
    from oracle import ranking as R
    from deeplearningrecommendationsystem_b200 import model as M, ops
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = torch.Generator().manual_seed(5)

    def dev_ms(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for nu, ni, W, k, tag in ((943, 1682, 64, 1682, "ML-100k full ranking (scripts/mf.py:81)"),
                              (65536, 16384, 64, 100, "synthetic 65536 users x 16384 items, top-100")):
        U = (torch.randn(nu, W, generator=g) * 0.3).to(dev)
        V = (torch.randn(ni, W, generator=g) * 0.3).to(dev)
        ms = dev_ms(lambda: ops.mf_rank(U, V, k, check=False), args.steps, max(args.warmup, 3))
        ms_lib = dev_ms(lambda: torch.topk(U @ V.T, k, dim=1), max(args.steps // 4, 2), 2)
        ns = min(nu, 256)
        Un, Vn = U[:ns].cpu().numpy(), V.cpu().numpy()
        t0 = time.perf_counter()
        sc = (Un @ Vn.T).astype(np.float32)
        for u in range(ns):
            R.rank_desc(sc[u], k)
        cpu = ns / (time.perf_counter() - t0)
        print(json.dumps({"metric": "ranked users/sec", "value": nu / (ms / 1e3), "unit": "users/s", "n_gpus": 1, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": f"rank/mf: {tag}", "users": nu, "items": ni, "width": W, "k": k},
                          "library_route_ms": ms_lib, "gpu_launches": args.steps,
                          "cpu_baseline": {"value": cpu, "unit": "users/s", "cores": 1, "kind": "port", "sample": f"{ns} users, numpy"}}))

    nu, ni = 943, 1682
    torch.manual_seed(0)
    m = M.DeepFM(nu, ni, [256, 128, 1], 128).to(dev).eval()          # scripts/deepfm.py:52
    df = catalogue_frame(g, nu, ni, shuffle=False)
    m.recommendation(32, df[df["user_id"] < 32], ni)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = m.recommendation(nu, df, ni)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    ns = 24
    t0 = time.perf_counter()
    with torch.no_grad():
        for u in range(ns):                                            # the reference's loop, model/deepfm.py:85-95, on this GPU
            rows = torch.Tensor(df[df["user_id"] == u].values).to(dev)
            torch.topk(m(rows), ni, dim=0).indices.view(1, -1).tolist()[0]
    loop = ns / (time.perf_counter() - t0)
    print(json.dumps({"metric": "ranked users/sec", "value": nu / sec, "unit": "users/s", "n_gpus": 1, "steps": 1, "warmup": 1,
                      "ms_per_step": sec * 1e3, "higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "rank/deepfm: DeepFM.recommendation(943, user_item, 1682), wall clock incl. host grouping",
                                 "rows": int(len(df)), "shape": list(out.shape)},
                      "per_user_loop_users_per_s": loop, "per_user_loop_sample": f"{ns} users, same GPU, reference control flow"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dist", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--light", action="store_true", help="cap every cardinality at 2^17 rows (fits any GPU)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--graph", dest="graph", action="store_true", default=None,
                    help="replay the train step as a CUDA graph (default for the single-GPU c2 headline)")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="eager Trainer.train_loop instead of graph replay")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5", "rank"],
                    help="c2 = headline (BASELINE.json configs[1]); c3/c4/c5 = the other synthetic configs, one line per model")
    args = ap.parse_args()
    if args.graph is None:
        # single GPU: the end-to-end leg replays the step as a CUDA graph (graph.GraphedTrainStep) -- same kernels, but
        # the host cost per step drops to two launches, which makes that loop immune to host jitter.  The sharded
        # multi-GPU step has a host sync (all-to-all split sizes) and stays eager.
        args.graph = args.workload == "c2" and int(os.environ.get("WORLD_SIZE", "1")) == 1
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c2":
        run_gpu(args)
    elif args.workload == "rank":
        run_rank(args)
    else:
        run_other(args)


if __name__ == "__main__":
    main()
