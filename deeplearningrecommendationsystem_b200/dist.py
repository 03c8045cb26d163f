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
        if ops.PROFILE is not None and torch.cuda.is_available():
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


class PeerRowExchange(RowExchange):
    """RowExchange whose two heavy all-to-alls are fused into the kernels over NVLink peer memory.

    Every rank owns two symmetric (peer-mapped) buffers per row width: `block` (the rows its batch needs) and `grads`
    (the per-row gradients other ranks push to it).  Forward: the OWNER's gather kernel stores each requested row
    straight into the requester's `block` (rs_gather_rows_peer) -- no staging buffer, no NCCL copy.  Backward: the
    requester's segment-reduce kernel stores each reduced row gradient straight into the owner's `grads`
    (RS_UPD_GRAD with grad_routes).  A device-side barrier on the symmetric-memory signal pads orders the phases.
    Only the small id/count messages still go through NCCL.
    """

    def __init__(self, prims, group=None, grads_slack=2.0):
        super().__init__(prims, group)
        self.grads_slack = grads_slack
        self._bufs = {}

    def _buffers(self, width, cap_rows, device):
        import torch.distributed._symmetric_memory as symm
        key = (width, device)
        ent = self._bufs.get(key)
        if ent is None or ent["cap"] < cap_rows:
            group = self.group if self.group is not None else dist.group.WORLD
            gcap = int(cap_rows * self.grads_slack)
            block = symm.empty((cap_rows, width), dtype=torch.float32, device=device)
            grads = symm.empty((gcap, width), dtype=torch.float32, device=device)
            hb = symm.rendezvous(block, group.group_name)
            hg = symm.rendezvous(grads, group.group_name)
            ent = {"cap": cap_rows, "gcap": gcap, "block": block, "grads": grads, "hb": hb, "hg": hg}
            self._bufs[key] = ent
        return ent

    def _plan(self, keys, total_rows):
        N = self.world
        R = (total_rows + N - 1) // N
        okeys = (keys % N) * R + keys // N
        uniq, inverse, *rest = self.prims.unique(okeys, N * R)
        bounds = torch.arange(N + 1, device=keys.device, dtype=uniq.dtype) * R
        send_counts_t = torch.diff(torch.searchsorted(uniq, bounds))
        # every rank learns the whole (requester, owner) count matrix in ONE all-gather: that fixes all split sizes
        # and all destination offsets of both fused exchanges
        allc = torch.empty(N, N, dtype=send_counts_t.dtype, device=keys.device)
        dist.all_gather_into_tensor(allc, send_counts_t, group=self.group)
        counts = allc.tolist()                                        # the one host sync of the plan
        send_counts = counts[self.rank]
        recv_counts = [counts[r][self.rank] for r in range(N)]
        send_local = uniq % R
        recv_local = torch.empty(sum(recv_counts), dtype=torch.int64, device=keys.device)
        dist.all_to_all_single(recv_local, send_local, recv_counts, send_counts, group=self.group)
        plan = Plan(int(uniq.numel()), inverse, send_local, send_counts, recv_counts, recv_local, rest[0] if rest else None)
        plan.counts = counts
        return plan

    def fetch(self, plan, local_table):
        from . import ops
        N, me, W = self.world, self.rank, local_table.shape[1]
        cap = plan.local_ids.numel()                                  # a batch cannot need more distinct rows than lookups
        ent = self._buffers(W, cap, local_table.device)
        counts = plan.counts
        # rows requested by rank r start at sum(recv_counts[:r]) in my request list and belong at
        # sum(counts[r][:me]) in r's block
        starts, row0s = [0], []
        for r in range(N):
            starts.append(starts[-1] + counts[r][me])
            row0s.append(sum(counts[r][:me]))
        routes = ops.make_routes(starts, [ent["hb"].buffer_ptrs[r] for r in range(N)], row0s)
        ops.gather_rows_peer(local_table, plan.recv_local, routes)
        with _phase("exchange_barrier"):
            ent["hb"].barrier()
        return ent["block"][: plan.n_uniq]

    def grad_routes(self, plan, width, device):
        """Routes for the reduced row gradients of the fetched block: block rows of owner o go to o's `grads` buffer at
        the offset where this rank's rows start in o's request list."""
        from . import ops
        N, me = self.world, self.rank
        ent = self._bufs[(width, device)]
        counts = plan.counts
        need = max(sum(counts[r][o] for r in range(N)) for o in range(N))
        if need > ent["gcap"]:
            raise RuntimeError(f"PeerRowExchange: an owner receives {need} rows > capacity {ent['gcap']}; raise grads_slack")
        starts, row0s = [0], []
        for o in range(N):
            starts.append(starts[-1] + counts[me][o])
            row0s.append(sum(counts[r][o] for r in range(me)))
        return ops.make_routes(starts, [ent["hg"].buffer_ptrs[o] for o in range(N)], row0s), ent

    def finish_push(self, plan, ent):
        """After the routed segment-reduce: barrier, then this rank's received gradients (aligned with recv_local)."""
        with _phase("exchange_barrier"):
            ent["hg"].barrier()
        return ent["grads"][: plan.recv_local.numel()]


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


def allreduce_dense_grads(params, group=None):
    """Average the gradients of the replicated dense parameters with ONE all-reduce over a flat bucket."""
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
