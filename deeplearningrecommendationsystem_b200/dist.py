"""Multi-GPU partition of the hot path (one process per GPU, torch.distributed / NCCL over NVLink).

The reference is single-process; this is new design.  Two regimes:

* small tables + all dense parameters: replicated; gradients are averaged with one all-reduce
  (``allreduce_dense_grads``) -- classic data parallelism;
* large tables: ROW-SHARDED.  Global row r lives on rank ``r % N`` at local index ``r // N`` (uniform load even
  under Zipf ids).  A step exchanges three all-to-alls, all of them on the *deduplicated* rows of the local batch:
      forward   ids -> owners, rows -> requesters          (``RowExchange.plan`` / ``fetch``)
      backward  per-row reduced gradients -> owners        (``RowExchange.push_grads``)
  and the owner runs the deterministic sort / segment-reduce / row-update over what it received.
  Deduplicating before the exchange is what keeps NVLink traffic below HBM traffic: a batch of 65536 x 26
  Criteo-shaped lookups touches ~0.5 M distinct rows, not 1.7 M.

The index math is backend-agnostic (``prims`` supplies unique / gather) so the routing is tested on CPU with gloo
(tests/test_dist_cpu.py); the product path binds the CUDA kernels (``cuda_prims``).
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist


def _phase(name):
    """CUDA-event bracket used by bench.py's per-phase timing (no-op unless ops.PROFILE is set, or on CPU)."""
    import contextlib
    try:
        from . import ops
        if (ops.PROFILE is not None or ops.NVTX) and torch.cuda.is_available():
            return ops._timed(name)
    except Exception:
        pass
    return contextlib.nullcontext()


@dataclass
class Prims:
    unique: callable      # (keys (n,) int64, total_rows) -> (uniq ascending (nu,), inverse (n,) int64[, sort handle])
    gather: callable      # (table (R, W), idx (m,) int64) -> (m, W)


def cuda_prims():
    from . import ops

    def unique(keys, total_rows):
        segs = ops.dedup_sort(keys, 1, None, total_rows, max_width=1, reuse_workspace=False)
        # third value: the sort itself, which the backward re-uses for the per-row gradient reduce (Plan.segs)
        return segs.uniq().clone(), segs.inverse().long(), segs

    def gather(table, idx):
        return ops.gather_rows(ops.make_tables([table]), idx.view(-1, 1)).view(idx.numel(), table.shape[1])

    return Prims(unique, gather)


@dataclass
class Plan:
    n_uniq: int
    local_ids: torch.Tensor      # (n,) index of each lookup's row inside the fetched block (owner-grouped order)
    send_rows: torch.Tensor      # (nu,) owner-local row indices requested, grouped by owner rank
    send_counts: list
    recv_counts: list
    recv_local: torch.Tensor     # (m,) local row indices other ranks asked this rank for (grouped by requester)
    segs: object = None          # CUDA path: the stable sort / segments of the batch's keys (ops.Segments), else None


class RowExchange:
    _memo = {}    # (device, group) -> (key tensor, version, total_rows, Plan): the plan depends only on the keys

    def __init__(self, prims, group=None):
        self.prims, self.group = prims, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def plan_for(self, ids, offsets, total_rows):
        """plan() on keys = ids + offsets, memoised on the identity of `ids`: two tables indexed by the same id
        tensor with the same row offsets (the FM and the FFM model of one batch) exchange the same rows."""
        slot = (ids.device, id(self.group))
        hit = RowExchange._memo.get(slot)
        if hit is not None and hit[0] is ids and hit[1] == ids._version and hit[2] == total_rows and hit[3] is offsets:
            return hit[4]
        plan = self.plan((ids + offsets).reshape(-1), total_rows)
        RowExchange._memo[slot] = (ids, ids._version, total_rows, offsets, plan)
        return plan

    def local_rows(self, total_rows):
        return (total_rows - self.rank + self.world - 1) // self.world

    def plan(self, keys, total_rows):
        """keys: (n,) int64 GLOBAL row of every lookup of the local batch."""
        with _phase("exchange_plan"):
            return self._plan(keys, total_rows)

    def _plan(self, keys, total_rows):
        N = self.world
        R = (total_rows + N - 1) // N                       # local rows per rank (upper bound)
        # Sort key = owner * R + local index: ONE dedup then yields the unique rows already grouped by owner and
        # ascending inside each group, so `inverse` is directly the position inside the fetched block.
        okeys = (keys % N) * R + keys // N if N > 1 else keys
        uniq, inverse, *rest = self.prims.unique(okeys, N * R)
        bounds = torch.arange(N + 1, device=keys.device, dtype=uniq.dtype) * R
        send_counts_t = torch.diff(torch.searchsorted(uniq, bounds))
        recv_counts_t = torch.empty_like(send_counts_t)
        if N > 1:
            dist.all_to_all_single(recv_counts_t, send_counts_t, group=self.group)
        else:
            recv_counts_t.copy_(send_counts_t)
        send_counts, recv_counts = torch.stack([send_counts_t, recv_counts_t]).tolist()   # the one host sync of the plan
        send_local = uniq % R if N > 1 else uniq              # local index at the owner
        recv_local = torch.empty(sum(recv_counts), dtype=torch.int64, device=keys.device)
        if N > 1:
            dist.all_to_all_single(recv_local, send_local, recv_counts, send_counts, group=self.group)
        else:
            recv_local.copy_(send_local)
        return Plan(int(uniq.numel()), inverse, send_local, send_counts, recv_counts, recv_local, rest[0] if rest else None)

    def fetch(self, plan, local_table):
        """-> (n_uniq, W) block holding the rows this rank's batch needs, in the order `local_ids` indexes."""
        mine = self.prims.gather(local_table, plan.recv_local)
        out = torch.empty(plan.n_uniq, local_table.shape[1], dtype=local_table.dtype, device=local_table.device)
        with _phase("exchange_rows"):
            if self.world > 1:
                dist.all_to_all_single(out, mine, plan.send_counts, plan.recv_counts, group=self.group)
            else:
                out.copy_(mine)
        return out

    def push_grads(self, plan, grads):
        """grads (n_uniq, W) in fetched-block order -> (m, W) aligned with plan.recv_local on the owners."""
        out = torch.empty(plan.recv_local.numel(), grads.shape[1], dtype=grads.dtype, device=grads.device)
        with _phase("exchange_grads"):
            if self.world > 1:
                dist.all_to_all_single(out, grads.contiguous(), plan.recv_counts, plan.send_counts, group=self.group)
            else:
                out.copy_(grads)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# Device-side exchange: no host synchronisation anywhere in the step
class Fabric:
    """What the device-side exchange needs from the interconnect: symmetric (peer-mapped) buffers and a stream-ordered
    cross-rank barrier.  ``alloc`` must be called in the same order with the same sizes on every rank."""
    world, rank = 1, 0

    def alloc(self, shape, dtype, device):
        """-> (this rank's tensor, [device address of every rank's tensor, indexed by rank])"""
        raise NotImplementedError

    def barrier(self):
        raise NotImplementedError

    def max_over_ranks(self, value):
        return int(value)


class LocalFabric(Fabric):
    """world == 1: the same kernels and the same launch sequence, buffers are plain device memory."""

    def alloc(self, shape, dtype, device):
        t = torch.zeros(shape, dtype=dtype, device=device)
        return t, [t.data_ptr()]

    def barrier(self):
        pass


class SymmFabric(Fabric):
    """torch symmetric memory over NVLink / NVSwitch (plumbing: allocation, peer pointers and the signal-pad barrier
    kernel; every byte of payload is moved by this package's own kernels through those pointers)."""

    def __init__(self, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self._handles, self._hb = [], None

    def alloc(self, shape, dtype, device):
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(shape, dtype=dtype, device=device)
        h = symm.rendezvous(t, self.group.group_name)
        t.zero_()
        self._handles.append(h)
        if self._hb is None:
            self._hb = h
        return t, [int(h.buffer_ptrs[r]) for r in range(self.world)]

    def barrier(self):
        self._hb.barrier()

    def max_over_ranks(self, value):
        t = torch.tensor([int(value)], dtype=torch.int64, device=torch.cuda.current_device())
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    AR_MAX = 4096      # floats; larger dense buckets go through NCCL (NVLS), small ones are latency bound

    def allreduce_mean(self, grads):
        """Average a SMALL set of dense gradients (the FM/FFM/MF bias, a few floats) through a symmetric buffer: every rank
        copies its flat bucket into slot [rank] of every peer, barrier, fixed-order sum.  No NCCL call, so the sharded
        step stays one launch sequence.  Two buffers are used alternately, so ONE barrier per call is enough: nobody can
        overwrite the buffer a slow rank is still summing before that rank has passed the next call's barrier.
        Returns False (caller falls back to NCCL) for big buckets."""
        K = sum(g.numel() for g in grads)
        if K == 0:
            return True
        if K > self.AR_MAX or getattr(self, "_ar_broken", False):
            return False
        dev = grads[0].device
        if getattr(self, "_ar", None) is None:
            try:
                buf, _ = self.alloc((2, self.world, self.AR_MAX), torch.float32, dev)
                h = self._handles[-1]
                views = [h.get_buffer(r, (2, self.world, self.AR_MAX), torch.float32) for r in range(self.world)]
                self._ar, self._ar_flip = (buf, views), 0
                self.barrier()
            except Exception:                      # symmetric-memory API without get_buffer: use NCCL
                self._ar_broken = True
                return False
        buf, views = self._ar
        flip = self._ar_flip
        if not torch.cuda.is_current_stream_capturing():
            self._ar_flip ^= 1         # (a captured step always replays the buffer it was captured with: see below)
        flat = torch.cat([g.reshape(-1) for g in grads])
        for r in range(self.world):
            views[r][flip, self.rank, :K].copy_(flat)
        self.barrier()
        mean = buf[flip, :, :K].sum(0) / self.world
        off = 0
        for g in grads:
            g.copy_(mean[off:off + g.numel()].view_as(g))
            off += g.numel()
        if torch.cuda.is_current_stream_capturing():
            self.barrier()             # a replayed graph cannot alternate buffers: keep the second barrier there
        return True


class ThreadFabric(Fabric):
    """N virtual ranks as N python threads on ONE GPU (tests and the single-GPU verification leg of bench.py): the
    kernels only ever see pointers, so 'peer' buffers are simply the other threads' tensors.  Ranks take turns (one lock),
    every barrier drains the rank's stream before the next rank runs -- slow, but it executes exactly the kernels and
    the launch order of the multi-GPU step."""

    class Shared:
        def __init__(self, world):
            import threading
            self.world = world
            self.lock = threading.Lock()
            self.bar = threading.Barrier(world)
            self.allocs = {}

    def __init__(self, shared, rank):
        self.shared, self.world, self.rank = shared, shared.world, rank
        self._seq = 0

    def alloc(self, shape, dtype, device):
        key, self._seq = self._seq, self._seq + 1
        sh = self.shared
        if key not in sh.allocs:                 # the first rank to get here allocates for everyone (the lock is held)
            sh.allocs[key] = [torch.zeros(shape, dtype=dtype, device=device) for _ in range(self.world)]
        ts = sh.allocs[key]
        return ts[self.rank], [t.data_ptr() for t in ts]

    def barrier(self):
        torch.cuda.current_stream().synchronize()
        self.shared.lock.release()
        try:
            self.shared.bar.wait(timeout=120)
        finally:
            self.shared.lock.acquire()

    def allreduce_mean(self, grads):
        """average the given gradient tensors over the virtual ranks (rank order, so every rank gets the same bits)"""
        sh = self.shared
        sh.allocs.setdefault("ar", [None] * self.world)[self.rank] = [g.clone() for g in grads]
        self.barrier()
        for i, g in enumerate(grads):
            acc = sh.allocs["ar"][0][i].clone()
            for r in range(1, self.world):
                acc += sh.allocs["ar"][r][i]
            g.copy_(acc / self.world)
        self.barrier()

    @staticmethod
    def run(world, fn):
        """fn(fabric) on `world` threads; returns the list of results by rank (exceptions re-raised)."""
        import threading
        shared = ThreadFabric.Shared(world)
        out, err = [None] * world, [None] * world

        def body(r):
            fab = ThreadFabric(shared, r)
            try:
                with shared.lock:
                    s = torch.cuda.Stream()
                    with torch.cuda.stream(s):
                        out[r] = fn(fab)
                        s.synchronize()
            except BaseException as e:            # noqa: BLE001
                err[r] = e
                shared.bar.abort()
        ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for e in err:
            if e is not None and not isinstance(e, threading.BrokenBarrierError):
                raise e
        for e in err:
            if e is not None:
                raise e
        return out


class DevPlan:
    """One step's exchange plan; every count lives on the device.  Valid until the exchange plans the next batch."""

    def __init__(self, gen, n, segs, local_ids, recv_local, m_total):
        self.gen, self.n, self.segs, self.local_ids, self.recv_local, self.m_total = gen, n, segs, local_ids, recv_local, m_total
        self.osegs, self.osegs_ready = None, None
        self.ready = None            # event: the plan was built on a side stream (plan_for(on_side=True))

    @property
    def n_uniq(self):             # capacity stand-in: callers only use it to size things
        return self.n


class DeviceRowExchange:
    """Row exchange of a sharded table with the plan kept on the device (rs_shard_* kernels, csrc/shard.cu).

    Every rank owns symmetric buffers: `req` / `ctl` (request lists and counts, written by the requesters), and per row
    width a `block` (the rows its batch needs, written by the owners' rs_shard_serve) and a `grads` buffer (reduced row
    gradients, written by the requesters' segment-reduce through rs_routes).  Payload moves only inside this package's
    kernels over NVLink peer pointers; barriers are stream ordered; there is no NCCL call and no host read in a step, so
    the whole train step can be captured into a CUDA graph.

    Capacities are fixed at first use from the number of lookups per batch (agreed over the ranks): a batch can never
    need more distinct rows than lookups, and an owner is given room for recv_slack x that many incoming rows; an
    overflow sets bit 16 of the status word (ops.check_status raises)."""

    device_plan = True

    def __init__(self, fabric=None, recv_slack=1.25, direct=()):
        """direct: global-row ranges [(lo, hi), ...] that a forward kernel reads from the owners' shards by itself
        (nfield.FieldFFM over symmetric table shards); their rows are requested (the owner needs the list for its update)
        but flagged, and fetch(..., skip_direct=True) does not copy them."""
        self.direct = tuple((int(lo), int(hi)) for lo, hi in direct)
        if fabric is None:
            fabric = SymmFabric() if dist.is_initialized() and dist.get_world_size() > 1 else LocalFabric()
        self.fabric, self.world, self.rank, self.recv_slack = fabric, fabric.world, fabric.rank, recv_slack
        self._ctl = None
        self._bufs = {}
        self._gen, self._memo = 0, None
        self._side, self._unjoined = None, None
        self._derived = None

    def local_rows(self, total_rows):
        return (total_rows - self.rank + self.world - 1) // self.world

    # ---- buffers
    def _control(self, n, total_rows, device):
        from . import _lib, ops
        R = (total_rows + self.world - 1) // self.world
        if self._ctl is not None:
            c = self._ctl
            if n > c["cap_req"] or R != c["R"]:
                raise RuntimeError(f"DeviceRowExchange was sized for {c['cap_req']} lookups per batch over {c['R']} rows per rank; "
                                   f"got {n} lookups / {R} rows (build a new exchange for a different table or a larger batch)")
            return c
        cap_req = self.fabric.max_over_ranks(n)
        cap_recv = int(cap_req * self.recv_slack) + 64
        req, req_ptrs = self.fabric.alloc((self.world * cap_req,), torch.int32, device)
        ctl, ctl_ptrs = self.fabric.alloc((_lib.RS_SHARD_CTL_WORDS,), torch.int64, device)
        self._ctl = {"cap_req": cap_req, "cap_recv": cap_recv, "R": R, "req": req, "ctl": ctl,
                     "S": ops.make_shard(self.world, self.rank, R, cap_req, cap_recv, req_ptrs, ctl_ptrs, self.direct),
                     "recv_local": torch.zeros(cap_recv, dtype=torch.int64, device=device),
                     "recv_skip": torch.zeros(cap_recv, dtype=torch.uint8, device=device),
                     "m_total": torch.zeros(1, dtype=torch.int32, device=device)}
        self.fabric.barrier()                      # every rank's buffers exist (and are zeroed) before anyone writes to them
        return self._ctl

    def _buffers(self, table, device, need_block=True):
        """block / grads buffers of one table (keyed by the table, not just its width: the fetched block must survive
        until that table's backward, and a model may shard two tables of equal width).  need_block=False: the table's rows
        are never fetched (its forward reads the shards directly), only the gradient buffer is needed."""
        width, key = table.shape[1], (table.shape[1], table.data_ptr())
        ent = self._bufs.get(key)
        if ent is None:
            c = self._ctl
            block, block_ptrs = self.fabric.alloc((c["cap_req"], width), torch.float32, device) if need_block else (None, None)
            grads, grad_ptrs = self.fabric.alloc((c["cap_recv"], width), torch.float32, device)
            ent = self._bufs[key] = {"block": block, "block_ptrs": block_ptrs, "grads": grads, "grad_ptrs": grad_ptrs}
            self.fabric.barrier()
        elif need_block and ent["block"] is None:
            ent["block"], ent["block_ptrs"] = self.fabric.alloc((self._ctl["cap_req"], width), torch.float32, device)
            self.fabric.barrier()
        return ent

    # ---- forward
    def derived(self, ids, name, make):
        """A tensor derived from the batch's id matrix (e.g. the columns of the row-sharded fields), built once per batch and
        shared by every model stepping on that batch -- so that their plans and sorts are shared as well."""
        d = self._derived
        if d is None or d[0] is not ids or d[1] != ids._version:
            d = self._derived = (ids, ids._version, {})
        if name not in d[2]:
            d[2][name] = make()
        return d[2][name]

    def plan_for(self, ids, offsets, total_rows, on_side=False, prefetch_rows=None, wait=True):
        """ids (B, F) local ids, offsets: HOST list of the F row offsets.  Memoised on the identity of `ids` while it is
        the exchange's latest plan (the FM and the FFM model of one batch share it).
        on_side: build the plan on the exchange's side stream (the caller's forward does not need it: it overlaps);
        prefetch_rows: also run the owner-side sort there (training); wait=False: do not order the current stream after a
        plan that is still being built -- the caller does that later with wait_plan()."""
        m = self._memo
        if m is not None and m[0] is ids and m[1] == ids._version and m[2] == total_rows and m[3] == tuple(offsets):
            plan = m[4]
            if wait:
                self.wait_plan(plan)
            if prefetch_rows is not None and plan.osegs is None:
                self.wait_plan(plan)
                self.prefetch_owner_segments(plan, prefetch_rows)
            return plan
        cur = torch.cuda.current_stream(ids.device)
        if on_side:
            if self._side is None:
                self._side = torch.cuda.Stream(device=ids.device)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                with _phase("exchange_plan"):
                    plan = self._plan(ids, offsets, total_rows)
                if prefetch_rows is not None:
                    self.prefetch_owner_segments(plan, prefetch_rows)
                plan.ready = torch.cuda.Event()
                plan.ready.record(self._side)
        else:
            with _phase("exchange_plan"):
                plan = self._plan(ids, offsets, total_rows)
            if prefetch_rows is not None:
                self.prefetch_owner_segments(plan, prefetch_rows)
        self._memo = (ids, ids._version, total_rows, tuple(offsets), plan)
        return plan

    def wait_plan(self, plan):
        """order the current stream after a plan built on the side stream"""
        if plan is not None and plan.ready is not None:
            torch.cuda.current_stream().wait_event(plan.ready)

    def _plan(self, ids, offsets, total_rows):
        from . import ops
        F = ids.shape[1] if ids.dim() == 2 else 1
        n = ids.numel()
        c = self._control(n, total_rows, ids.device)
        if self._unjoined is not None:             # an unconsumed owner-side sort of the previous plan still reads recv_local
            torch.cuda.current_stream(ids.device).wait_event(self._unjoined)
            self._unjoined = None
        segs = ops.dedup_sort(ids, F, list(offsets) if F > 1 or offsets else None, total_rows, shard=(self.world, c["R"]))
        ops.shard_post(c["S"], segs)
        self.fabric.barrier()
        ops.shard_collect(c["S"], c["recv_local"], c["m_total"], c["recv_skip"])
        self._gen += 1
        return DevPlan(self._gen, n, segs, segs.inverse().long(), c["recv_local"], c["m_total"])

    def _check(self, plan):
        if plan.gen != self._gen:
            raise RuntimeError("stale exchange plan: the exchange has planned another batch since (one batch in flight per exchange)")

    def fetch(self, plan, local_table, skip_direct=False):
        """-> (cap, W) block; rows [0, n_uniq) hold the rows this rank's batch needs, in the order `local_ids` indexes.
        skip_direct: rows of the direct ranges are left out (the caller's forward kernel reads them from the shards)."""
        from . import ops
        self._check(plan)
        c, ent = self._ctl, self._buffers(local_table, local_table.device)
        ops.shard_serve(c["S"], local_table, plan.recv_local, ent["block_ptrs"], c["cap_req"],
                        skip=c["recv_skip"] if (skip_direct and self.direct) else None)
        with _phase("exchange_barrier"):
            self.fabric.barrier()
        return ent["block"]

    # ---- backward
    def prefetch_owner_segments(self, plan, local_rows):
        """The owner-side sort depends only on the plan: run it on a side stream under the forward kernels."""
        from . import ops
        if plan.osegs is not None:
            return
        dev = plan.recv_local.device
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        self._side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._side):
            plan.osegs = ops.dedup_sort(plan.recv_local, 1, None, local_rows, n_valid=plan.m_total)
            plan.osegs_ready = torch.cuda.Event()
            plan.osegs_ready.record()
        self._unjoined = plan.osegs_ready

    def owner_segments(self, plan, local_rows, width):
        from . import ops
        self._check(plan)
        if plan.osegs is None:
            plan.osegs = ops.dedup_sort(plan.recv_local, 1, None, local_rows, n_valid=plan.m_total)
        elif plan.osegs_ready is not None:
            torch.cuda.current_stream().wait_event(plan.osegs_ready)
            if self._unjoined is plan.osegs_ready:
                self._unjoined = None
        return ops.attach_partial(plan.osegs, width)

    def grad_routes(self, plan, table, device):
        """Routes of the reduced row gradients of the fetched block: block rows of owner o (a device-resident range) go to
        o's `grads` buffer at the offset o announced during rs_shard_collect."""
        from . import _lib, ops
        self._check(plan)
        c, ent = self._ctl, self._buffers(table, device, need_block=False)
        base = c["ctl"].data_ptr()
        return ops.make_routes(None, ent["grad_ptrs"], None, dyn_start=base + 8 * _lib.RS_CTL_SEND_START,
                               dyn_row0=base + 8 * _lib.RS_CTL_G0_IN, cap_rows=c["cap_recv"], self_index=self.rank), ent

    def finish_push(self, plan, ent):
        with _phase("exchange_barrier"):
            self.fabric.barrier()
        return ent["grads"]


def shard_rows(global_table, rank, world):
    """The rows of a replicated/global table that `rank` owns (r % world == rank), as a contiguous copy."""
    return global_table[rank::world].contiguous()


def unshard_rows(shards):
    """Inverse of shard_rows for tests: list of per-rank shards -> global table."""
    world = len(shards)
    total = sum(s.shape[0] for s in shards)
    out = torch.empty(total, shards[0].shape[1], dtype=shards[0].dtype, device=shards[0].device)
    for r, s in enumerate(shards):
        out[r::world] = s
    return out


def allreduce_dense_grads(params, group=None, fabric=None):
    """Average the gradients of the replicated dense parameters with ONE all-reduce over a flat bucket."""
    if fabric is not None and hasattr(fabric, "allreduce_mean"):          # symmetric-memory bucket / virtual ranks
        if fabric.allreduce_mean([p.grad for p in params if p.grad is not None]) is not False:
            return
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
