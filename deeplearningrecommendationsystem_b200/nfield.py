"""N-field generalisations of the reference's FM / FFM interaction models for the synthetic Criteo-shaped
configs (BASELINE.json configs[1]: 26 sparse fields, D = 16, B = 65536).

The reference modules are hard-wired to the six MovieLens features; these keep the same arithmetic (FM second
order model/deepfm.py:71-77, FFM pairs model/ffm.py:61-82, sigmoid head, BCELoss outside) for F id-fields.  All F
tables live in ONE concatenated (total_rows, W) tensor; ids are int64 (B, F), local to each field.

With ``sharded=True`` (multi-GPU) the concatenated table is ROW-SHARDED over the process group (global row r on rank
r % N) and every step runs the deduplicated ids / rows / row-gradients all-to-alls of ``dist.RowExchange``; the
result equals the single-GPU step on the concatenated global batch (gradients are averaged over ranks).

Two update modes:
  fused=True   (default) the table is not an autograd leaf.  forward launches the fused lookup+interaction kernel
               and keeps the Jacobian rows; backward only records dL/dcross; ``FusedRowOptimizer.step()`` runs the
               sort / segment-reduce / row-update kernels.
  fused=False  the table is a normal dense-gradient nn.Parameter (the reference's semantics, any torch optimizer);
               the gradient is produced by the same deterministic segment-reduce.
"""
import math

import torch
from torch import nn

from . import ops


def _xavier_concat(cards, width, row_dim, device=None, seed=None):
    """xavier_normal_ per field table, as the reference does for each nn.Embedding (model/deepfm.py:34-41).
    Filled slice by slice in place so a 56 GB table never needs a second copy."""
    out = torch.empty(sum(cards), width, dtype=torch.float32, device=device)
    g = None
    if seed is not None:
        g = torch.Generator(device=out.device).manual_seed(seed)
    r0 = 0
    for c in cards:
        out[r0:r0 + c].normal_(0.0, math.sqrt(2.0 / (c + row_dim)), generator=g)
        r0 += c
    return out


class _CrossFn(torch.autograd.Function):
    """Routes dL/dcross into the owning module; the table itself is updated by the fused optimizer."""

    @staticmethod
    def forward(ctx, anchor, cross, owner, token):
        ctx.owner, ctx.token = owner, token
        return cross.clone()

    @staticmethod
    def backward(ctx, g):
        ctx.owner._record(ctx.token, g.contiguous())
        return torch.zeros(1, device=g.device), None, None, None


class _DenseGradFn(torch.autograd.Function):
    """Dense-gradient mode: d loss / d table via the deterministic segment-reduce (RS_UPD_GRAD)."""

    @staticmethod
    def forward(ctx, weight, cross, owner, ids, stash):
        ctx.owner = owner
        ctx.save_for_backward(ids, stash)
        ctx.shape = weight.shape
        return cross.clone()

    @staticmethod
    def backward(ctx, g):
        ids, stash = ctx.saved_tensors
        o = ctx.owner
        segs = ops.dedup_sort(ids, o.F, o.offsets_host, o.total_rows, max_width=o.width)
        grad = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        ops.segment_update(segs, ops.RS_UPD_GRAD, o.width, o.F, stash=stash, scale=g.contiguous(), dense_grad=grad)
        return grad, None, None, None, None


class _FieldModel(nn.Module):
    def __init__(self, cardinalities, width, row_dim, fused=True, seed=None, device=None, sharded=False, group=None):
        super().__init__()
        self.cards = [int(c) for c in cardinalities]
        self.F, self.width, self.fused = len(self.cards), width, fused
        offs = [0]
        for c in self.cards:
            offs.append(offs[-1] + c)
        self.offsets_host, self.total_rows = offs[:-1], offs[-1]
        self.sharded, self.exchange = sharded, None
        if sharded:
            if not fused:
                raise ValueError("sharded tables are updated by the fused row optimizer (fused=True)")
            from . import dist as rsdist
            self.exchange = rsdist.RowExchange(rsdist.cuda_prims(), group)
            rows = self.exchange.local_rows(self.total_rows)
            std = math.sqrt(2.0 / (self.total_rows / self.F + row_dim))
            w = torch.empty(rows, width, dtype=torch.float32, device=device)
            g = torch.Generator(device=w.device).manual_seed((seed or 0) * 1000 + self.exchange.rank)
            self.weight = nn.Parameter(w.normal_(0.0, std, generator=g), requires_grad=False)
            self.register_buffer("offsets_dev", torch.tensor(self.offsets_host, dtype=torch.int64, device=device), persistent=False)
        else:
            self.weight = nn.Parameter(_xavier_concat(self.cards, width, row_dim, device, seed), requires_grad=not fused)
        self.bias = nn.Parameter(torch.zeros(1, device=device))
        self._pending, self._anchor, self._token = {}, None, 0
        self.adam_m = self.adam_v = None

    _offset_cache = {}

    def _offsets_key(self):
        """one shared device tensor per (device, offsets): models with equal field layouts share exchange plans"""
        key = (self.weight.device, tuple(self.offsets_host))
        t = _FieldModel._offset_cache.get(key)
        if t is None:
            t = torch.tensor(self.offsets_host, dtype=torch.int64, device=self.weight.device)
            _FieldModel._offset_cache[key] = t
        return t

    def load_global(self, global_weight):
        """sharded mode: take this rank's rows (r % N == rank) of a full (total_rows, W) table."""
        from . import dist as rsdist
        self.weight.data.copy_(rsdist.shard_rows(global_weight.to(self.weight.device), self.exchange.rank, self.exchange.world))

    def tables(self):
        return ops.tables_from_concat(self.weight.data, self.offsets_host, self.cards)

    # --- fused-mode bookkeeping
    def _record(self, token, g):
        if token in self._pending:
            self._pending[token]["g"] = g

    def clear_pending(self):
        self._pending.clear()

    def _row_update(self, opt, segs, F, **src):
        """Apply the optimizer to the rows of self.weight named by `segs` (gradient source in **src)."""
        if opt.kind == "sgd":
            ops.segment_update(segs, ops.RS_UPD_SGD, self.width, F, table=self.weight.data, lr=opt.lr, wd=opt.weight_decay, **src)
        else:
            if self.adam_m is None:
                self.adam_m = torch.zeros_like(self.weight.data)
                self.adam_v = torch.zeros_like(self.weight.data)
            ops.segment_update(segs, ops.RS_UPD_ADAM, self.width, F, table=self.weight.data, m=self.adam_m, v=self.adam_v,
                               lr=opt.lr, wd=opt.weight_decay, betas=opt.betas, eps=opt.eps, step=opt.step_count, **src)

    def apply_pending(self, opt):
        for rec in self._pending.values():
            if "g" not in rec:
                continue
            if not self.sharded:
                segs = ops.dedup_sort(rec["ids"], self.F, self.offsets_host, self.total_rows, max_width=self.width)
                self._row_update(opt, segs, self.F, stash=rec["stash"], scale=rec["g"])
                continue
            # row-sharded: reduce the batch's gradients per fetched row, send them to the owners, update there
            plan, ex = rec["plan"], self.exchange
            segs = ops.dedup_sort(plan.local_ids, 1, None, plan.n_uniq, max_width=self.width)   # memoised per plan
            # every fetched row has at least one lookup in this batch, so RS_UPD_GRAD writes every row of the block
            block_grad = torch.empty(plan.n_uniq, self.width, dtype=torch.float32, device=rec["g"].device)
            ops.segment_update(segs, ops.RS_UPD_GRAD, self.width, self.F, stash=rec["stash"], scale=rec["g"] * (1.0 / ex.world),
                               dense_grad=block_grad)
            recv = ex.push_grads(plan, block_grad)
            if recv.shape[0]:
                osegs = ops.dedup_sort(plan.recv_local, 1, None, self.weight.shape[0], max_width=self.width)
                self._row_update(opt, osegs, 1, dense=recv)
        self._pending.clear()

    def _interact(self, T, ids, want_stash):
        raise NotImplementedError

    def logit(self, ids):
        if ids.dim() != 2 or ids.shape[1] != self.F:
            raise ValueError(f"ids must be (B, {self.F})")
        train = torch.is_grad_enabled()
        rec = {}
        if self.sharded:
            plan = self.exchange.plan_for(ids, self._offsets_key(), self.total_rows)
            block = self.exchange.fetch(plan, self.weight.data)
            ids = plan.local_ids.view(ids.shape)
            cross, stash = self._interact(ops.make_tables([block] * self.F), ids, want_stash=train)
            rec = {"plan": plan}
        else:
            cross, stash = self._interact(self.tables(), ids, want_stash=train)
        if train and self.fused:
            if self._anchor is None or self._anchor.device != ids.device:
                self._anchor = torch.zeros(1, device=ids.device, requires_grad=True)
            self._token += 1
            self._pending[self._token] = {"ids": ids, "stash": stash, **rec}
            cross = _CrossFn.apply(self._anchor, cross, self, self._token)
        elif train:
            cross = _DenseGradFn.apply(self.weight, cross, self, ids, stash)
        return cross + self.bias if self.use_bias else cross

    use_bias = True

    def forward(self, ids):
        return torch.sigmoid(self.logit(ids)).unsqueeze(1)


class FieldFM(_FieldModel):
    """sigmoid(b + 0.5 * sum_d[(sum_f e_f)^2 - sum_f e_f^2]) over F id-fields, D-dim rows."""

    def __init__(self, cardinalities, embedding_dim, fused=True, seed=None, device=None, sharded=False, group=None):
        super().__init__(cardinalities, embedding_dim, embedding_dim, fused, seed, device, sharded, group)

    def _interact(self, T, ids, want_stash):
        out = ops.fields_fwd(T, ids.shape[0], ids.device, ids=ids, cross=True, stash=want_stash)
        return out["cross"], out.get("stash")


class FieldFFM(_FieldModel):
    """sigmoid(b + sum_{i<j} <v_{i,j}, v_{j,i}>): feature i's table row is (F, D), slot j aimed at field j."""

    def __init__(self, cardinalities, num_vector, fused=True, seed=None, device=None, sharded=False, group=None):
        F = len(cardinalities)
        super().__init__(cardinalities, F * num_vector, num_vector, fused, seed, device, sharded, group)
        self.D = num_vector

    def _interact(self, T, ids, want_stash):
        return ops.ffm_fwd(T, ids, self.D, want_stash=want_stash)


class FieldMF(_FieldModel):
    """sigmoid(<U[u], V[i]>) with the user and item tables in one concatenated (optionally row-sharded) buffer --
    the large-table form of reference model/mf.py:11-26 (same forward signature and 1-D output) for the synthetic
    100 M-row configs.  d dot / d U[u] = V[i] is exactly the FM Jacobian S - e for two fields, so the fused lookup
    kernel's stash and the segment-reduce/update path are shared with FieldFM."""

    use_bias = False

    def __init__(self, num_users, num_items, embedding_size, fused=True, seed=None, device=None, sharded=False, group=None):
        super().__init__([num_users, num_items], embedding_size, embedding_size, fused, seed, device, sharded, group)
        self.bias.requires_grad_(False)

    def _interact(self, T, ids, want_stash):
        out = ops.fields_fwd(T, ids.shape[0], ids.device, ids=ids, dot2=True, stash=want_stash)
        return out["dot2"], out.get("stash")

    def forward(self, user_indices, item_indices):
        ids = torch.stack([user_indices, item_indices], dim=1)
        return torch.sigmoid(self.logit(ids))                       # (B,)
