"""Attention pooling (AFM), target attention (DIN / DIEN) and the GRU recurrence of the drop-in modules."""
import torch


class _AFMPool(torch.autograd.Function):
    """rs_afm_fwd / rs_afm_bwd: the pair tensor stays in shared memory, parameter gradients are fixed-order sums."""

    @staticmethod
    def forward(ctx, E, W, b, h):
        from . import ops
        need = any(t.requires_grad for t in (E, W, b, h))
        pooled, attw = ops.afm_fwd(E, W.detach(), b.detach(), h.detach(), want_attw=need)
        if need:
            ctx.save_for_backward(E, W, b, h, attw)
        return pooled

    @staticmethod
    def backward(ctx, g):
        from . import ops
        E, W, b, h, attw = ctx.saved_tensors
        dE, dW, db, dh = ops.afm_bwd(E, W.detach(), b.detach(), h.detach(), attw, g.contiguous())
        return dE, dW, db, dh.view_as(h)


def afm_pool(E, W, b, h):
    """Attention pooling over the pairwise Hadamard products of E (B, F, D) -> (B, D)   reference model/afm.py:55-65."""
    return _AFMPool.apply(E.contiguous(), W, b, h)


class _DINAttention(torch.autograd.Function):
    """rs_din_fwd / rs_din_bwd on rows (B, L+1, D) = [history | target].  Large batches run the hidden layers on the
    tensor cores (rs_din_fwd_tc); that forward stashes its ReLU outputs and rs_din_bwd_tc consumes them, while the
    CUDA-core kernel pair recomputes instead."""

    @staticmethod
    def forward(ctx, rows, pool, W0, b0, W1, b1, W2, b2):
        from . import ops
        ws = (W0, b0, W1, b1, W2, b2)
        need = any(ctx.needs_input_grad)
        out, _, stash = ops.din_fwd(rows, [t.detach() for t in ws], pool, want_stash=need)
        ctx.pool, ctx.tc = pool, stash is not None
        ctx.save_for_backward(rows, *ws, *(stash or ()))
        return out

    @staticmethod
    def backward(ctx, g):
        from . import ops
        rows, *rest = ctx.saved_tensors
        ws, stash = rest[:6], rest[6:]
        if ctx.tc:
            d_rows, dws = ops.din_bwd_tc(rows, [t.detach() for t in ws], ctx.pool, g.contiguous(), stash)
        else:
            d_rows, dws = ops.din_bwd(rows, [t.detach() for t in ws], ctx.pool, g.contiguous())
        return (d_rows, None, *dws)


class _DINAttentionTC(torch.autograd.Function):
    """The attention unit as tensor-core GEMMs over all B*L (sample, position) rows: rs_gemm_nt_3xtf32 for the layer
    products (bias / per-sample target term / ReLU / ReLU-mask fused in the epilogue) and rs_gemm_tn_3xtf32 for the
    weight gradients (reductions over B*L).  The concat is never built: W0 z = (Wa+Wb) h + (Wc-Wb) t.  Softmax and the
    (B, L, D) weighting are small elementwise ops."""

    @staticmethod
    def forward(ctx, rows, pool, W0, b0, W1, b1, W2, b2):
        from . import ops
        B, L1, D = rows.shape
        L = L1 - 1
        Wab = (W0[:, :D] + W0[:, D:2 * D]).contiguous()
        Wcb = (W0[:, 2 * D:] - W0[:, D:2 * D]).contiguous()
        hist_view = (L, L1 * D, D)                                            # row r -> rows[r // L, r % L]
        c = ops.gemm_nt(rows[:, L], Wcb, bias=b0, M=B, a_rows=(1, L1 * D, D))                      # (B, H1) target part
        a1 = ops.gemm_nt(rows, Wab, rowbias=c, rb_group=L, relu=True, M=B * L, a_rows=hist_view)  # (B*L, H1)
        a2 = ops.gemm_nt(a1, W1, bias=b1, relu=True)                                              # (B*L, H2)
        s = ops.gemm_nt(a2, W2, bias=b2).view(B, L)
        w = torch.softmax(s, dim=1)
        h = rows[:, :L]
        out = torch.einsum("bl,bld->bd", w, h) if pool else h * w.unsqueeze(-1)
        ctx.pool = pool
        ctx.save_for_backward(rows, a1, a2, w, Wab, Wcb, W1, W2)
        return out

    @staticmethod
    def backward(ctx, g):
        from . import ops
        rows, a1, a2, w, Wab, Wcb, W1, W2 = ctx.saved_tensors
        B, L1, D = rows.shape
        L = L1 - 1
        h, t = rows[:, :L], rows[:, L]
        if ctx.pool:
            dw = torch.einsum("bld,bd->bl", h, g)
            dh = w.unsqueeze(-1) * g.unsqueeze(1)
        else:
            dw = (h * g).sum(dim=-1)
            dh = w.unsqueeze(-1) * g
        ds = (w * (dw - (w * dw).sum(dim=1, keepdim=True))).reshape(B * L, 1)
        dW2 = ops.gemm_tn(ds, a2)                                             # (1, H2)
        db2 = ds.sum().view(1)
        da2 = torch.where(a2 > 0, ds * W2, torch.zeros((), device=a2.device))  # (B*L, H2)
        dW1 = ops.gemm_tn(da2, a1)                                            # (H2, H1)
        db1 = da2.sum(dim=0)
        da1 = ops.gemm_nt(da2, W1.t().contiguous(), mask=a1)                  # (da2 W1) * [a1 > 0]
        dWab = ops.gemm_tn(da1, h.reshape(B * L, D))                          # (H1, D)
        db0 = da1.sum(dim=0)
        s1 = da1.view(B, L, -1).sum(dim=1)                                    # (B, H1)
        dWt = ops.gemm_tn(s1, t.contiguous())                                 # (H1, D)
        dW0 = torch.cat([dWab, dWab - dWt, dWt], dim=1)
        dh = dh + ops.gemm_nt(da1, Wab.t().contiguous()).view(B, L, D)
        dt = ops.gemm_nt(s1, Wcb.t().contiguous())                            # (B, D)
        return torch.cat([dh, dt.unsqueeze(1)], dim=1), None, dW0, db0, dW1, db1, dW2, db2


def din_attention(rows, unit, pool, impl="fused"):
    """softmax_L(MLP([h, h-t, t])) applied to the history                               reference model/din.py:39-47.
    rows (B, L+1, D): gathered history rows followed by the target row; unit = Sequential(Linear, ReLU, Linear, ReLU,
    Linear).  pool=True -> (B, D) weighted sum; pool=False -> (B, L, D) scaled history (model/dien.py:33-37).
    impl "fused" (default): rs_din_fwd / rs_din_bwd -- for B*L >= 8192 rows the tcgen05 kernels of din_tc.cu (both hidden
    layers fused per 128-row tile, 3xTF32), otherwise the CUDA-core kernel pair; "tc": the same unit written as separate
    rs_gemm_nt_3xtf32 / rs_gemm_tn_3xtf32 calls per layer (kept as a parity cross-check; slower, intermediates in HBM)."""
    l0, l1, l2 = unit[0], unit[2], unit[4]
    fn = _DINAttentionTC if impl == "tc" else _DINAttention
    return fn.apply(rows.contiguous(), pool, l0.weight, l0.bias, l1.weight, l1.bias, l2.weight, l2.bias)


class _GRURecurrence(torch.autograd.Function):
    """L sequential GRU steps on precomputed input projections (rs_gru_fwd / rs_gru_bwd).  Returns h_L (B, H)."""

    @staticmethod
    def forward(ctx, gi, w_hh, b_hh):
        from . import ops
        need = gi.requires_grad or w_hh.requires_grad or b_hh.requires_grad
        h_all, gates = ops.gru_fwd(gi, w_hh.detach(), b_hh.detach(), want_gates=need)
        if need:
            ctx.save_for_backward(w_hh, h_all, gates)
        return h_all[:, -1].clone()

    @staticmethod
    def backward(ctx, g_last):
        from . import ops
        w_hh, h_all, gates = ctx.saved_tensors
        B, L, H = h_all.shape
        d_gi, d_gh = ops.gru_bwd(w_hh.detach(), h_all, gates, g_h_last=g_last.contiguous())
        h_prev = torch.cat([torch.zeros(B, 1, H, device=h_all.device), h_all[:, :-1]], dim=1)
        flat = d_gh.view(B * L, 3 * H)
        return d_gi, flat.t() @ h_prev.reshape(B * L, H), flat.sum(dim=0)


def gru_last_hidden(x, gru):
    """hidden[-1] of a single-layer batch_first GRU started from zeros                  reference model/dien.py:61-64.
    Input projection = one library GEMM; the recurrence and its BPTT run in the hand-written kernels."""
    H = gru.hidden_size
    if gru.num_layers != 1 or gru.bidirectional or not gru.batch_first or H not in (8, 16, 32, 64):
        raise NotImplementedError("gru_last_hidden: single-layer batch_first GRU with hidden size in {8,16,32,64}")
    gi = torch.nn.functional.linear(x, gru.weight_ih_l0, gru.bias_ih_l0)
    return _GRURecurrence.apply(gi, gru.weight_hh_l0, gru.bias_hh_l0)
