"""Parity of every C-ABI kernel against the CPU oracle on seeded inputs (run on the B200: -m gpu).

Bars (north star): bit-exact for gather / dedup / index work; 1e-5 relative (fp32) for interactions, gradients
and updated rows.
"""
import numpy as np
import pytest
import torch

from oracle import index as oidx, interactions as OI, optim as ooptim

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6


def _ops():
    from deeplearningrecommendationsystem_b200 import ops
    return ops


def close(got, want, rtol=RTOL, atol=ATOL, msg=""):
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().cpu().numpy(), rtol=rtol, atol=atol, err_msg=msg)


def make_fields(F, D, rows, B, seed, zipf=False):
    g = torch.Generator().manual_seed(seed)
    tabs = [torch.randn(r, D, generator=g) * 0.3 for r in rows]
    if zipf:  # heavy duplicates: most lookups hit the first few rows
        ids = torch.stack([(torch.rand(B, generator=g) ** 4 * r).long().clamp_(max=r - 1) for r in rows], dim=1)
    else:
        ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], dim=1)
    return tabs, ids


# ------------------------------------------------------------------ gather (bit-exact)
@pytest.mark.parametrize("F,W,B", [(1, 16, 1000), (3, 64, 257), (26, 16, 513), (2, 1, 300), (4, 10, 77), (2, 256, 65)])
def test_gather_bit_exact(F, W, B):
    ops = _ops()
    rows = [7 + 13 * f for f in range(F)]
    tabs, ids = make_fields(F, W, rows, B, seed=F * 100 + W)
    T = ops.make_tables([t.cuda() for t in tabs])
    keep = [t.cuda() for t in tabs]
    T = ops.make_tables(keep)
    out = ops.gather_rows(T, ids.cuda())
    want = np.stack([oidx.gather(tabs[f].numpy(), ids[:, f].numpy()) for f in range(F)], axis=1)
    assert np.array_equal(out.cpu().numpy(), want)
    ops.check_status()


def test_gather_out_of_range_sets_status():
    ops = _ops()
    tab = torch.randn(5, 8).cuda()
    ids = torch.tensor([[0], [5], [2]]).cuda()
    ops.gather_rows(ops.make_tables([tab]), ids)
    with pytest.raises(IndexError):
        ops.check_status()
    ops.check_status()  # flag was cleared


# ------------------------------------------------------------------ dedup (bit-exact)
@pytest.mark.parametrize("n,F,card", [(1, 1, 5), (64, 1, 3), (1000, 1, 50), (4096, 4, 17), (26 * 700, 26, 1000), (65 * 129, 1, 1)])
def test_dedup_matches_unique(n, F, card):
    ops = _ops()
    g = torch.Generator().manual_seed(n + F)
    ids = torch.randint(0, card, (n,), generator=g)
    offs = [f * card for f in range(F)]
    segs = ops.dedup_sort(ids.cuda(), F, offs, F * card, reuse_workspace=False)
    keys = ids.numpy() + np.tile(np.asarray(offs), n // F)
    u, inv, cnt = oidx.dedup(keys)
    assert segs.n_uniq == len(u)
    assert np.array_equal(segs.uniq().cpu().numpy(), u)
    assert np.array_equal(segs.inverse().cpu().numpy().astype(np.int64), inv)
    assert np.array_equal(segs.counts().cpu().numpy().astype(np.int64), cnt)
    assert np.array_equal(segs.sorted_pos().cpu().numpy().astype(np.int64), oidx.stable_order(keys))


@pytest.mark.parametrize("n,total", [(1, 7), (33, 2 ** 31 + 5), (4095, 300), (4096, 2 ** 9), (4097, 2 ** 9 + 1), (70001, 2 ** 18), (1703936, 33762577),
                                     (3000017, 2 ** 32 - 1)])
def test_radix_sort_is_stable_and_equals_library_sort(n, total):
    """the hand-written LSD radix sort + prefix sums (sort.cu, RS_SORT=own) against numpy's stable argsort, and -- array
    by array -- against the default dedup driven by cub: tile edges, 1..4 digit passes, heavy duplicates, 32-bit keys"""
    import os
    ops = _ops()
    g = torch.Generator().manual_seed(n)
    ids = (torch.rand(n, generator=g, dtype=torch.float64) ** 3 * total).long().clamp_(max=total - 1)
    ids[::7] = total - 1
    cu = ids.cuda()
    lib = ops.dedup_sort(cu, 1, None, total, reuse_workspace=False)
    os.environ["RS_SORT"] = "own"
    try:
        own = ops.dedup_sort(cu, 1, None, total, reuse_workspace=False)
    finally:
        del os.environ["RS_SORT"]
    order = np.argsort(ids.numpy(), kind="stable")
    assert np.array_equal(own.sorted_pos().cpu().numpy().astype(np.int64), order)
    assert own.n_uniq == lib.n_uniq == len(np.unique(ids.numpy()))
    for name in ("sorted_pos", "inverse", "counts", "uniq"):
        assert torch.equal(getattr(own, name)(), getattr(lib, name)()), name
    ops.check_status()


# ------------------------------------------------------------------ segment reduce + update
@pytest.mark.parametrize("W", [1, 4, 8, 16, 64, 128, 416, 10, 1248])
@pytest.mark.parametrize("hot", [False, True])
def test_segment_grad_and_sgd(W, hot):
    ops = _ops()
    g = torch.Generator().manual_seed(W + hot)
    rows, n = 300, 5000
    ids = (torch.rand(n, generator=g) ** (6 if hot else 1) * rows).long().clamp_(max=rows - 1)
    if hot:
        ids[::3] = 7                      # one segment of > 1600 lookups -> many chunks
    G = torch.randn(n, W, generator=g)
    table = torch.randn(rows, W, generator=g)
    segs = ops.dedup_sort(ids.cuda(), 1, None, rows, max_width=W, reuse_workspace=False)
    dense = torch.zeros(rows, W).cuda()
    ops.segment_update(segs, ops.RS_UPD_GRAD, W, 1, dense=G.cuda(), dense_grad=dense)
    want = torch.zeros(rows, W).index_add_(0, ids, G)
    # a re-associated fp32 sum of n terms differs from the sequential one by a few ulp of the L1 mass of the
    # terms, so the absolute floor scales with sum|G| (3 eps); elsewhere the bar is 1e-5 relative
    atol = max(1e-6, 2e-7 * float(torch.zeros(rows, W).index_add_(0, ids, G.abs()).max()))
    close(dense, want, rtol=1e-5, atol=atol)
    # determinism: bitwise identical on a second run
    dense2 = torch.zeros(rows, W).cuda()
    ops.segment_update(segs, ops.RS_UPD_GRAD, W, 1, dense=G.cuda(), dense_grad=dense2)
    assert torch.equal(dense, dense2)
    tc = table.clone().cuda()
    ops.segment_update(segs, ops.RS_UPD_SGD, W, 1, dense=G.cuda(), table=tc, lr=0.05)
    close(tc, ooptim.sgd_rows(table.clone(), ids, G, 0.05), rtol=1e-5, atol=atol)


@pytest.mark.parametrize("W", [16, 416])           # 16: chunk kernel, 416: TMA streaming kernel (reads the lookup records)
def test_relabelled_segments_equal_a_second_sort(W):
    """rs_segments_relabel: the sort of a batch's global rows, renamed to block rows, drives the per-row gradient
    reduce exactly like sorting the block-local ids again (what the row-sharded backward used to do)."""
    ops = _ops()
    g = torch.Generator().manual_seed(W)
    total, n = 10 ** 7, 20000
    keys = (torch.rand(n, generator=g) ** 4 * total).long().clamp_(max=total - 1)
    keys[::5] = 123456                                     # a long multi-chunk segment
    stash = torch.randn(n, W, generator=g).cuda()
    scale = torch.randn(n, generator=g).cuda()
    a = ops.dedup_sort(keys.cuda(), 1, None, total, max_width=1, reuse_workspace=False)
    uniq, inverse = a.uniq().clone(), a.inverse().long()
    nu = uniq.numel()
    assert torch.equal(uniq.cpu(), torch.unique(keys))
    a = ops.block_segments(a, nu, W)
    b = ops.dedup_sort(inverse, 1, None, nu, max_width=W, reuse_workspace=False)
    ga, gb = torch.zeros(nu, W).cuda(), torch.zeros(nu, W).cuda()
    ops.segment_update(a, ops.RS_UPD_GRAD, W, 1, stash=stash, scale=scale, dense_grad=ga)
    ops.segment_update(b, ops.RS_UPD_GRAD, W, 1, stash=stash, scale=scale, dense_grad=gb)
    assert torch.equal(ga, gb)
    G = (stash * scale[:, None]).cpu()
    want = torch.zeros(nu, W).index_add_(0, inverse.cpu(), G)
    atol = max(1e-6, 2e-7 * float(torch.zeros(nu, W).index_add_(0, inverse.cpu(), G.abs()).max()))   # see test_segment_grad_and_sgd
    close(ga, want, rtol=1e-5, atol=atol)
    ops.check_status()


def test_segment_short_segments_bit_exact_vs_sequential():
    """segments <= RS_CHUNK are summed in ascending position, the CPU index_add order -> bit-identical."""
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    rows, n, W = 2000, 6000, 16
    ids = torch.randint(0, rows, (n,), generator=g)
    G = torch.randn(n, W, generator=g)
    segs = ops.dedup_sort(ids.cuda(), 1, None, rows, max_width=W, reuse_workspace=False)
    dense = torch.zeros(rows, W).cuda()
    ops.segment_update(segs, ops.RS_UPD_GRAD, W, 1, dense=G.cuda(), dense_grad=dense)
    want = torch.zeros(rows, W)
    for p in range(n):                         # strictly sequential fp32 adds
        want[ids[p]] += G[p]
    assert torch.equal(dense.cpu(), want)


@pytest.mark.parametrize("W", [8, 128])          # 8: lane-group kernel, 128: TMA streaming kernel
def test_segment_scaled_stash_and_adam(W):
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    F, B, card = 3, 400, 40
    ids = torch.randint(0, card, (B, F), generator=g)
    offs = [0, card, 2 * card]
    stash = torch.randn(B, F, W, generator=g)
    scale = torch.randn(B, generator=g)
    vscale = torch.randn(B, W, generator=g)
    extra = torch.randn(B, F, W, generator=g)
    keys = (ids + torch.tensor(offs)).reshape(-1)
    table = torch.randn(F * card, W, generator=g)
    segs = ops.dedup_sort(ids.cuda(), F, offs, F * card, max_width=W, reuse_workspace=False)
    # scalar scale + dense term, SGD with weight decay
    G = (scale[:, None, None] * stash + extra).reshape(-1, W)
    tc = table.clone().cuda()
    ops.segment_update(segs, ops.RS_UPD_SGD, W, F, stash=stash.cuda(), scale=scale.cuda(), dense=extra.cuda(), table=tc, lr=0.1, wd=0.01)
    uniq, gs = ooptim.segment_sum_rows(keys, G)
    want = table.clone()
    want[uniq] -= 0.1 * (gs + 0.01 * want[uniq])
    close(tc, want)
    # vector scale, lazy Adam, two steps
    G2 = (vscale[:, None, :] * stash).reshape(-1, W)
    tc, m, v = table.clone().cuda(), torch.zeros_like(table).cuda(), torch.zeros_like(table).cuda()
    wt, wm, wv = table.clone(), torch.zeros_like(table), torch.zeros_like(table)
    for step in (1, 2):
        ops.segment_update(segs, ops.RS_UPD_ADAM, W, F, stash=stash.cuda(), scale=vscale.cuda(), table=tc, m=m, v=v, lr=1e-2, step=step)
        ooptim.adam_rows(wt, wm, wv, keys, G2, step, lr=1e-2)
    # Adam's step is lr*m/(sqrt(v)+eps): rows whose summed gradient nearly cancels move in a direction set by
    # rounding noise, so the table gets the same 2%-of-lr absolute floor as the model tests
    close(tc, wt, rtol=1e-5, atol=2e-4)
    close(m, wm, rtol=1e-5, atol=2e-6)
    close(v, wv, rtol=1e-5, atol=2e-6)


def test_adam_dense_matches_torch_adam():
    ops = _ops()
    torch.manual_seed(0)
    p0, grads = torch.randn(1000), [torch.randn(1000) for _ in range(3)]
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=1e-5)
    p, m, v = p0.clone().cuda(), torch.zeros(1000).cuda(), torch.zeros(1000).cuda()
    for s, g in enumerate(grads, 1):
        ref.grad = g.clone()
        opt.step()
        ops.adam_dense(p, g.cuda(), m, v, s, lr=1e-3, wd=1e-5)
    close(p, ref.data, rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------ fused lookup + interaction
@pytest.mark.parametrize("F,D,B", [(26, 16, 1000), (6, 128, 300), (39, 32, 257), (2, 64, 500), (2, 256, 130), (6, 8, 96), (5, 4, 64)])
@pytest.mark.parametrize("src", ["ids", "dense"])
def test_fields_fwd(F, D, B, src):
    ops = _ops()
    rows = [11 + 7 * f for f in range(F)]
    tabs, ids = make_fields(F, D, rows, B, seed=F * D)
    E = torch.stack([tabs[f][ids[:, f]] for f in range(F)], dim=1)
    keep = [t.cuda() for t in tabs]
    two = F == 2
    if src == "ids":
        out = ops.fields_fwd(ops.make_tables(keep), B, "cuda", ids=ids.cuda(), cross=True, bi=True, pairs=True, concat=True,
                             stash=True, dot2=two, had2=two)
    else:
        out = ops.fields_fwd(ops.dummy_tables(F, D), B, "cuda", dense_in=E.cuda(), cross=True, bi=True, pairs=True, concat=True,
                             stash=True, dot2=two, had2=two)
    assert torch.equal(out["concat"].cpu(), E.flatten(1))                   # gather part is a pure copy
    close(out["cross"], OI.fm_second_order(E), rtol=1e-5, atol=2e-5)
    close(out["bi"], OI.bi_interaction(E), rtol=1e-5, atol=1e-5)
    close(out["pairs"], OI.inner_products(E), rtol=1e-5, atol=2e-6)
    close(out["stash"], E.sum(1, keepdim=True) - E, rtol=1e-5, atol=2e-6)
    if two:
        close(out["dot2"], (E[:, 0] * E[:, 1]).sum(1), rtol=1e-5, atol=2e-6)
        assert torch.equal(out["had2"].cpu(), E[:, 0] * E[:, 1])
    ops.check_status()


@pytest.mark.parametrize("F,D,B", [(26, 16, 300), (6, 128, 100), (39, 32, 65), (2, 64, 200), (6, 8, 96)])
def test_fields_bwd(F, D, B):
    ops = _ops()
    rows = [11 + 7 * f for f in range(F)]
    tabs, ids = make_fields(F, D, rows, B, seed=F + D)
    g = torch.Generator().manual_seed(5)
    E = torch.stack([tabs[f][ids[:, f]] for f in range(F)], dim=1).requires_grad_(True)
    P = F * (F - 1) // 2
    gc, gb, gp, gcat = torch.randn(B, generator=g), torch.randn(B, D, generator=g), torch.randn(B, P, generator=g), torch.randn(B, F * D, generator=g)
    loss = (OI.fm_second_order(E) * gc).sum() + (OI.bi_interaction(E) * gb).sum() + (OI.inner_products(E) * gp).sum() + (E.flatten(1) * gcat).sum()
    two = F == 2
    if two:
        gd, gh = torch.randn(B, generator=g), torch.randn(B, D, generator=g)
        loss = loss + ((E[:, 0] * E[:, 1]).sum(1) * gd).sum() + (E[:, 0] * E[:, 1] * gh).sum()
    (want,) = torch.autograd.grad(loss, E)
    kw = dict(g_cross=gc.cuda(), g_bi=gb.cuda(), g_pairs=gp.cuda(), g_concat=gcat.cuda())
    if two:
        kw.update(g_dot2=gd.cuda(), g_had2=gh.cuda())
    keep = [t.cuda() for t in tabs]
    dE1 = ops.fields_bwd(ops.make_tables(keep), B, "cuda", ids=ids.cuda(), **kw)
    dE2 = ops.fields_bwd(ops.dummy_tables(F, D), B, "cuda", dense_in=E.detach().cuda(), **kw)
    close(dE1, want, rtol=1e-5, atol=2e-5)
    assert torch.equal(dE1, dE2)


# ------------------------------------------------------------------ FFM
@pytest.mark.parametrize("F,D,B", [(26, 16, 300), (4, 8, 1000), (39, 32, 20), (3, 4, 129), (8, 64, 50), (6, 128, 40)])
def test_ffm_fwd_and_stash(F, D, B):
    ops = _ops()
    rows = [5 + 3 * f for f in range(F)]
    tabs, ids = make_fields(F, F * D, rows, B, seed=F * 3 + D)
    T = torch.stack([tabs[f][ids[:, f]] for f in range(F)], dim=1).view(B, F, F, D).requires_grad_(True)
    want = OI.ffm_cross(T)
    (jac,) = torch.autograd.grad(want.sum(), T)
    keep = [t.cuda() for t in tabs]
    cross, stash = ops.ffm_fwd(ops.make_tables(keep), ids.cuda(), D, want_stash=True)
    close(cross, want.detach(), rtol=1e-5, atol=2e-5)
    assert torch.equal(stash.cpu().view(B, F, F, D), jac)                  # transposed copy with zero diagonal
    cross2, none = ops.ffm_fwd(ops.make_tables(keep), ids.cuda(), D, want_stash=False)
    assert none is None and torch.equal(cross, cross2)
    ops.check_status()


def test_ffm_dense():
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    B, F, NF, D = 70, 6, 2, 8
    fo = [0, 0, 0, 1, 0, 1]
    T = torch.randn(B, F, NF, D, generator=g, requires_grad=True)
    gc = torch.randn(B, generator=g)
    want = OI.ffm_cross(T, fo)
    (dT,) = torch.autograd.grad((want * gc).sum(), T)
    close(ops.ffm_dense_fwd(T.detach().cuda(), fo), want.detach(), rtol=1e-5, atol=2e-6)
    close(ops.ffm_dense_bwd(T.detach().cuda(), gc.cuda(), fo), dT, rtol=1e-5, atol=2e-6)


# ------------------------------------------------------------------ MovieLens feature-vector front end
def test_xembed_fwd_bwd():
    ops = _ops()
    from helpers import feature_matrix
    g = torch.Generator().manual_seed(4)
    B, D = 300, 16
    x = feature_matrix(g, B, 50, 60)
    tabs = [torch.randn(r, D, generator=g) for r in (50, 60, 1, 2, 21, 19)]
    slots = [(0, 1, 0, tabs[0].cuda()), (1, 1, 0, tabs[1].cuda()), (2, 1, 1, tabs[2].cuda()), (3, 2, 1, tabs[3].cuda()),
             (5, 21, 1, tabs[4].cuda()), (26, 19, 1, tabs[5].cuda()), (2, 1, 2, None)]
    S = ops.make_xslots(slots, D, 45)
    E = ops.xembed_fwd(S, x.cuda())
    want = torch.stack([tabs[0][x[:, 0].long()], tabs[1][x[:, 1].long()], x[:, 2:3] @ tabs[2], x[:, 3:5] @ tabs[3],
                        x[:, 5:26] @ tabs[4], x[:, 26:45] @ tabs[5], x[:, 2:3].expand(-1, D)], dim=1)
    close(E, want, rtol=1e-6, atol=1e-6)
    dE = torch.randn(B, 7, D, generator=g)
    dW = ops.xembed_bag_bwd(S, x.cuda(), dE.cuda(), slots)
    for t, (c0, nc) in {2: (2, 1), 3: (3, 2), 4: (5, 21), 5: (26, 19)}.items():
        close(dW[t], x[:, c0:c0 + nc].t() @ dE[:, t], rtol=1e-5, atol=1e-5)
    assert dW[0] is None and dW[6] is None
    assert torch.equal(ops.xcol_to_ids(x.cuda(), 1).cpu(), x[:, 1].long())


def test_sigmoid_bce():
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    z = (torch.randn(5000, generator=g) * 3).requires_grad_(True)
    z.data[0], z.data[1] = 200.0, -200.0              # exercise the -100 clamp
    y = (torch.rand(5000, generator=g) < 0.3).float()
    p = torch.sigmoid(z)
    loss = torch.nn.BCELoss()(p, y)
    pred, l, gz = ops.sigmoid_bce(z.detach().cuda(), y.cuda())
    close(pred, p.detach(), rtol=1e-6, atol=1e-7)
    close(l, OI.bce(p.detach(), y), rtol=1e-5)
    close(l, loss.detach(), rtol=1e-5)
    z.data[0], z.data[1] = 20.0, -20.0                # away from the clamp the analytic gradient is (p-y)/B
    pred, l, gz = ops.sigmoid_bce(z.detach().cuda(), y.cuda())
    (want,) = torch.autograd.grad(torch.nn.BCELoss()(torch.sigmoid(z), y), z)
    close(gz, want, rtol=1e-4, atol=1e-9)


# ------------------------------------------------------------------ AFM attention pooling
@pytest.mark.parametrize("F,D,A,B", [(6, 16, 8, 200), (6, 128, 64, 70), (39, 32, 64, 24), (5, 8, 40, 64), (3, 4, 1, 33),
                                     (7, 16, 32, 150), (4, 64, 32, 60), (6, 32, 32, 77), (9, 16, 64, 600)])
def test_afm_pool_fwd_bwd(F, D, A, B):
    ops = _ops()
    g = torch.Generator().manual_seed(F * D + A)
    E = (torch.randn(B, F, D, generator=g) * 0.5).requires_grad_(True)
    W = (torch.randn(D, A, generator=g) * 0.3).requires_grad_(True)
    b = torch.randn(A, generator=g).requires_grad_(True)
    h = torch.randn(A, 1, generator=g).requires_grad_(True)
    gp = torch.randn(B, D, generator=g)
    want = OI.afm_pool(E, W, b, h)
    gE, gW, gb, gh = torch.autograd.grad((want * gp).sum(), [E, W, b, h])
    pooled, attw = ops.afm_fwd(E.detach().cuda(), W.detach().cuda(), b.detach().cuda(), h.detach().cuda())
    # every output is a softmax-weighted sum of D- or A-term dot products: 1e-5 relative, with the absolute floor
    # tied to the tensor's own scale (elements far below the scale carry the rounding of the large terms)
    def scaled(got, ref):
        close(got, ref, rtol=1e-5, atol=max(1e-6, 1e-5 * float(ref.abs().max())))
    scaled(pooled, want.detach())
    dE, dW, db, dh = ops.afm_bwd(E.detach().cuda(), W.detach().cuda(), b.detach().cuda(), h.detach().cuda(), attw, gp.cuda())
    scaled(dE, gE)
    scaled(dW, gW)
    scaled(db, gb)
    scaled(dh, gh.view(-1))
    dE2, dW2, _, _ = ops.afm_bwd(E.detach().cuda(), W.detach().cuda(), b.detach().cuda(), h.detach().cuda(), attw, gp.cuda())
    assert torch.equal(dE, dE2) and torch.equal(dW, dW2)            # deterministic


@pytest.mark.parametrize("F,D,A,B", [(39, 32, 64, 700), (17, 16, 32, 1500), (20, 32, 128, 333)])
def test_afm_forward_tensor_cores(F, D, A, B):
    """batches of >= 2 x SMs samples take the tcgen05 projection (afm_tc.cu): same outputs as the CUDA-core kernel
    (forced with RS_AFM_NO_TC) and as the oracle; the unchanged backward consumes its attention weights."""
    import os
    ops = _ops()
    g = torch.Generator().manual_seed(F + B)
    E = (torch.randn(B, F, D, generator=g) * 0.5).requires_grad_(True)
    W = (torch.randn(D, A, generator=g) * 0.3).requires_grad_(True)
    b = torch.randn(A, generator=g).requires_grad_(True)
    h = torch.randn(A, 1, generator=g).requires_grad_(True)
    gp = torch.randn(B, D, generator=g)
    want = OI.afm_pool(E, W, b, h)
    gE, gW, gb, gh = torch.autograd.grad((want * gp).sum(), [E, W, b, h])
    cu = [t.detach().cuda() for t in (E, W, b, h)]
    pooled, attw = ops.afm_fwd(*cu)
    os.environ["RS_AFM_NO_TC"] = "1"
    try:
        pooled_cc, attw_cc = ops.afm_fwd(*cu)
    finally:
        del os.environ["RS_AFM_NO_TC"]
    assert not torch.equal(attw, attw_cc)           # two different kernels really ran
    close(attw, attw_cc.cpu(), rtol=1e-5, atol=1e-5 * float(attw_cc.max()))
    close(pooled, pooled_cc.cpu(), rtol=1e-5, atol=1e-5 * float(pooled_cc.abs().max()))
    close(pooled, want.detach(), rtol=1e-5, atol=max(1e-6, 1e-5 * float(want.detach().abs().max())))
    assert abs(float(attw.sum(1).mean()) - 1.0) < 1e-5
    pooled2, attw2 = ops.afm_fwd(*cu)
    assert torch.equal(pooled, pooled2) and torch.equal(attw, attw2)
    no_w, _ = ops.afm_fwd(*cu, want_attw=False)
    assert torch.equal(no_w, pooled)
    # B*P*A = tens of millions of ReLU inputs: a few sit within rounding of zero and may fall on either side in two fp32
    # evaluations, which moves the gradient of the two embeddings of that pair by a finite amount (see the DIN test).
    # Both backward implementations -- tcgen05 (what "auto" picks here) and CUDA cores -- against the oracle:
    for impl in ("auto", "cuda_cores"):
        n0 = ops.launches()
        dE, dW, db, dh = ops.afm_bwd(*cu, attw, gp.cuda(), impl=impl)
        assert ops.launches() - n0 == (3 if impl == "auto" else 1)          # chain + dE + dW kernels vs the single one
        err = (dE.cpu() - gE).abs()
        bad = int((err > 1e-5 * gE.abs() + 2e-5 * float(gE.abs().max())).flatten(1).any(dim=1).sum())
        assert bad <= 3 + B // 200, f"{impl}: {bad} samples differ"
        assert float(err.max()) <= 5e-2 * float(gE.abs().max()), impl
        for got, ref, name in ((dW, gW, "dW"), (db, gb, "db"), (dh, gh.view(-1), "dh")):
            assert got.shape == ref.shape, (impl, name)
            assert float((got.cpu() - ref).norm()) <= 1e-3 * float(ref.norm()), (impl, name)
        again = ops.afm_bwd(*cu, attw, gp.cuda(), impl=impl)
        assert all(torch.equal(x, y_) for x, y_ in zip((dE, dW, db, dh), again)), f"{impl}: not deterministic"
    # ... and the tcgen05 backward itself is checked EXACTLY (north star: 1e-5 everywhere): with the ReLU masks fixed to
    # the bits its chain kernel produced, the gradient is a smooth function of the inputs, and a float64 evaluation of it
    # (from the same attention weights the forward handed over) must agree element by element.
    dE, dW, db, dh, masks = ops.afm_bwd(*cu, attw, gp.cuda(), impl="auto", return_masks=True)
    i, j = OI.pair_index(F)
    E64, W64, b64, h64, g64 = [t.detach().double() for t in (E, W, b, h.view(-1), gp)]
    w64, m64 = attw.double().cpu(), masks.double().cpu()
    P = E64[:, i] * E64[:, j]                                           # (B, P, D)
    z = P @ W64 + b64
    assert float(((z > 0).double() - m64).abs().mean()) < 1e-4           # the masks are the ReLU's, up to kinks at |z| ~ 0
    assert float(torch.cat([z[(z > 0).double() != m64].abs().flatten(), torch.zeros(1, dtype=torch.float64)]).max()) < 1e-5
    dwp = (P * g64.unsqueeze(1)).sum(-1)                                # d loss / d w_p
    ds = w64 * (dwp - (w64 * dwp).sum(1, keepdim=True))
    dz = ds.unsqueeze(-1) * h64 * m64                                   # (B, P, A)
    dP = dz @ W64.t() + w64.unsqueeze(-1) * g64.unsqueeze(1)
    want_dE = torch.zeros_like(E64)
    want_dE.index_add_(1, i, dP * E64[:, j])
    want_dE.index_add_(1, j, dP * E64[:, i])
    want = {"dE": want_dE, "dW": torch.einsum("bpd,bpa->da", P, dz), "db": dz.sum((0, 1)), "dh": torch.einsum("bp,bpa->a", ds, m64 * z)}
    for name, got in (("dE", dE), ("dW", dW), ("db", db), ("dh", dh)):
        ref = want[name]
        close(got.double(), ref, rtol=1e-5, atol=1e-5 * float(ref.abs().max()), msg=name + " (fixed masks, f64)")


# ------------------------------------------------------------------ GRU recurrence
@pytest.mark.parametrize("H,B,L", [(16, 50, 7), (64, 33, 100), (8, 17, 3), (32, 16, 20)])
def test_gru_recurrence_fwd_bwd(H, B, L):
    ops = _ops()
    g = torch.Generator().manual_seed(H + L)
    D = H
    x = (torch.randn(B, L, D, generator=g) * 0.5).requires_grad_(True)
    k = 1.0 / H ** 0.5
    w_ih, w_hh = [((torch.rand(3 * H, n, generator=g) * 2 - 1) * k).requires_grad_(True) for n in (D, H)]
    b_ih, b_hh = [((torch.rand(3 * H, generator=g) * 2 - 1) * k).requires_grad_(True) for _ in range(2)]
    gl = torch.randn(B, H, generator=g)
    want = OI.gru(x, w_ih, w_hh, b_ih, b_hh)
    grads = torch.autograd.grad((want * gl).sum(), [x, w_ih, w_hh, b_ih, b_hh])
    # the oracle restatement itself equals torch.nn.GRU
    ref = torch.nn.GRU(D, H, batch_first=True)
    with torch.no_grad():
        ref.weight_ih_l0.copy_(w_ih), ref.weight_hh_l0.copy_(w_hh), ref.bias_ih_l0.copy_(b_ih), ref.bias_hh_l0.copy_(b_hh)
    np.testing.assert_allclose(ref(x)[1][-1].detach().numpy(), want.detach().numpy(), rtol=1e-5, atol=1e-6)
    gi = (x.detach() @ w_ih.detach().t() + b_ih.detach()).cuda()
    h_all, gates = ops.gru_fwd(gi, w_hh.detach().cuda(), b_hh.detach().cuda())
    close(h_all[:, -1], want.detach(), rtol=1e-5, atol=1e-6)
    d_gi, d_gh = ops.gru_bwd(w_hh.detach().cuda(), h_all, gates, g_h_last=gl.cuda())
    d_gi, d_gh = d_gi.cpu(), d_gh.cpu()
    h_prev = torch.cat([torch.zeros(B, 1, H), h_all.cpu()[:, :-1]], dim=1)
    got = [d_gi @ w_ih.detach(), d_gi.reshape(-1, 3 * H).t() @ x.detach().reshape(-1, D),
           d_gh.reshape(-1, 3 * H).t() @ h_prev.reshape(-1, H), d_gi.sum((0, 1)), d_gh.sum((0, 1))]
    for a, w_, name in zip(got, grads, ["dx", "dW_ih", "dW_hh", "db_ih", "db_hh"]):
        np.testing.assert_allclose(a.numpy(), w_.numpy(), rtol=1e-5, atol=2e-6 * max(1.0, float(w_.abs().max())), err_msg=name)


# ------------------------------------------------------------------ DIN target attention
@pytest.mark.parametrize("D,H1,H2,L,B,pool", [(16, 128, 64, 7, 90, True), (64, 128, 64, 100, 20, True), (16, 64, 32, 7, 90, False),
                                             (64, 64, 32, 100, 11, False), (32, 128, 64, 33, 40, True), (16, 128, 64, 1, 5, True)])
def test_din_attention_fwd_bwd(D, H1, H2, L, B, pool):
    ops = _ops()
    g = torch.Generator().manual_seed(D + L + H1)
    rows = (torch.randn(B, L + 1, D, generator=g) * 0.5).requires_grad_(True)
    lin = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5).requires_grad_(True)   # noqa: E731
    vec = lambda o, i: ((torch.rand(o, generator=g) * 2 - 1) / i ** 0.5).requires_grad_(True)      # noqa: E731
    ws = [lin(H1, 3 * D), vec(H1, 3 * D), lin(H2, H1), vec(H2, H1), lin(1, H2), vec(1, H2)]
    h, t = rows[:, :-1], rows[:, -1]
    w = OI.din_attention(h, t, [(ws[0], ws[1]), (ws[2], ws[3]), (ws[4], ws[5])])
    want = (h * w.unsqueeze(-1)).sum(1) if pool else h * w.unsqueeze(-1)
    gup = torch.randn(*want.shape, generator=g)
    grads = torch.autograd.grad((want * gup).sum(), [rows] + ws)
    cw = [t_.detach().cuda() for t_ in ws]
    out, attw = ops.din_fwd(rows.detach().cuda(), cw, pool, want_attw=True)

    def scaled(got, ref, name=""):
        close(got, ref, rtol=1e-5, atol=1e-5 * max(1e-3, float(ref.abs().max())), msg=name)
    scaled(attw, w.detach(), "attw")
    scaled(out, want.detach(), "out")
    # the tensor-core forward (what "auto" picks for large batches) against the same oracle, with and without attw
    out_tc, attw_tc = ops.din_fwd(rows.detach().cuda(), cw, pool, want_attw=True, impl="tc")
    scaled(attw_tc, w.detach(), "attw (tc)")
    scaled(out_tc, want.detach(), "out (tc)")
    out_tc2, _ = ops.din_fwd(rows.detach().cuda(), cw, pool, want_attw=False, impl="tc")
    assert torch.equal(out_tc, out_tc2)
    d_rows, dws = ops.din_bwd(rows.detach().cuda(), cw, pool, gup.cuda())
    scaled(d_rows, grads[0], "d_rows")
    for got, ref, name in zip(dws[:5], grads[1:6], ["dW0", "db0", "dW1", "db1", "dW2"]):
        scaled(got, ref, name)
    # softmax is shift invariant, so d loss / d b2 == 0 analytically: both sides are pure rounding noise
    assert float(dws[5].abs().max()) <= 1e-5 * max(1e-3, float(grads[5].abs().max()), float(grads[4].abs().max()))
    d2, dws2 = ops.din_bwd(rows.detach().cuda(), cw, pool, gup.cuda())
    assert torch.equal(d_rows, d2) and all(torch.equal(a, b_) for a, b_ in zip(dws, dws2))   # deterministic
    # tensor-core backward from the tensor-core forward's stash, same oracle gradients
    _, _, stash = ops.din_fwd(rows.detach().cuda(), cw, pool, impl="tc", want_stash=True)
    d_tc, dws_tc = ops.din_bwd_tc(rows.detach().cuda(), cw, pool, gup.cuda(), stash)
    scaled(d_tc, grads[0], "d_rows (tc)")
    for got, ref, name in zip(dws_tc[:5], grads[1:6], ["dW0", "db0", "dW1", "db1", "dW2"]):
        assert got.shape == ref.shape, name
        scaled(got, ref, name + " (tc)")
    assert float(dws_tc[5].abs().max()) <= 1e-5 * max(1e-3, float(grads[5].abs().max()), float(grads[4].abs().max()))
    d_tc2, dws_tc2 = ops.din_bwd_tc(rows.detach().cuda(), cw, pool, gup.cuda(), stash)
    assert torch.equal(d_tc, d_tc2) and all(torch.equal(a, b_) for a, b_ in zip(dws_tc, dws_tc2))


@pytest.mark.parametrize("D,H1,H2,L,B,pool", [(16, 64, 64, 5, 40, True), (32, 128, 32, 9, 30, False), (64, 64, 64, 1, 300, True)])
def test_din_tc_shapes_the_cuda_core_kernels_do_not_build(D, H1, H2, L, B, pool):
    """(H1, H2) in {(64,64), (128,32)} exist only on the tensor-core path: forward and backward against the oracle"""
    ops = _ops()
    g = torch.Generator().manual_seed(D + H1 + H2)
    rows = (torch.randn(B, L + 1, D, generator=g) * 0.5).requires_grad_(True)
    lin = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5).requires_grad_(True)   # noqa: E731
    vec = lambda o, i: ((torch.rand(o, generator=g) * 2 - 1) / i ** 0.5).requires_grad_(True)      # noqa: E731
    ws = [lin(H1, 3 * D), vec(H1, 3 * D), lin(H2, H1), vec(H2, H1), lin(1, H2), vec(1, H2)]
    h, t = rows[:, :-1], rows[:, -1]
    w = OI.din_attention(h, t, [(ws[0], ws[1]), (ws[2], ws[3]), (ws[4], ws[5])])
    want = (h * w.unsqueeze(-1)).sum(1) if pool else h * w.unsqueeze(-1)
    gup = torch.randn(*want.shape, generator=g)
    grads = torch.autograd.grad((want * gup).sum(), [rows] + ws)
    cw = [t_.detach().cuda() for t_ in ws]
    out, attw, stash = ops.din_fwd(rows.detach().cuda(), cw, pool, want_attw=True, impl="tc", want_stash=True)

    def scaled(got, ref, name=""):
        close(got, ref, rtol=1e-5, atol=1e-5 * max(1e-3, float(ref.abs().max())), msg=name)
    scaled(attw, w.detach(), "attw")
    scaled(out, want.detach(), "out")
    d_rows, dws = ops.din_bwd_tc(rows.detach().cuda(), cw, pool, gup.cuda(), stash)
    scaled(d_rows, grads[0], "d_rows")
    for got, ref, name in zip(dws[:5], grads[1:6], ["dW0", "db0", "dW1", "db1", "dW2"]):
        scaled(got, ref, name)
    with pytest.raises(RuntimeError):
        ops.din_fwd(rows.detach().cuda(), cw, pool, impl="fused")


@pytest.mark.parametrize("D,H1,H2,L,B,pool", [(64, 128, 64, 100, 700, True), (32, 64, 32, 50, 1000, False), (16, 128, 64, 3, 9000, True)])
def test_din_tc_forward_many_tiles(D, H1, H2, L, B, pool):
    """more tiles than SMs (persistent loop, buffer/barrier parity across tiles, ragged last tile); tc == fused kernel"""
    ops = _ops()
    g = torch.Generator().manual_seed(B)
    rows = (torch.randn(B, L + 1, D, generator=g) * 0.5).cuda()
    lin = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5).cuda()   # noqa: E731
    vec = lambda o, i: ((torch.rand(o, generator=g) * 2 - 1) / i ** 0.5).cuda()      # noqa: E731
    cw = [lin(H1, 3 * D), vec(H1, 3 * D), lin(H2, H1), vec(H2, H1), lin(1, H2), vec(1, H2)]
    out_f, attw_f = ops.din_fwd(rows, cw, pool, want_attw=True, impl="fused")
    out_t, attw_t = ops.din_fwd(rows, cw, pool, want_attw=True, impl="tc")
    close(attw_t, attw_f.cpu(), rtol=1e-5, atol=1e-5 * float(attw_f.abs().max()), msg="attw")
    close(out_t, out_f.cpu(), rtol=1e-5, atol=1e-5 * float(out_f.abs().max()), msg="out")
    out_t2, attw_t2 = ops.din_fwd(rows, cw, pool, want_attw=True, impl="tc")
    assert torch.equal(out_t, out_t2) and torch.equal(attw_t, attw_t2)       # deterministic
    gup = torch.randn(*out_f.shape, generator=g).cuda()
    d_f, dws_f = ops.din_bwd(rows, cw, pool, gup)
    _, _, stash = ops.din_fwd(rows, cw, pool, impl="tc", want_stash=True)
    d_t, dws_t = ops.din_bwd_tc(rows, cw, pool, gup, stash)
    # ReLU is not differentiable at 0: a pre-activation within rounding of zero can land on either side in two fp32
    # implementations, which changes that one (b, l) row's gradient by a finite amount.  13 M activations here, so a
    # handful of rows may differ; everything else must agree to 1e-5, and the outliers stay small.
    err = (d_t - d_f).abs()
    tol = 1e-5 * float(d_f.abs().max()) + 1e-5 * d_f.abs()
    bad_rows = int((err > tol).any(dim=2).sum())
    assert bad_rows <= max(2, int(2e-4 * B * L)), f"{bad_rows} history rows differ"
    assert float(err.max()) <= 1e-2 * float(d_f.abs().max())
    # Weight gradients sum those rows, so the same few kinks move them by more than 1e-5 between ANY two fp32
    # evaluations; against the fused kernel only the aggregate can be bounded ...
    for a, b_, name in zip(dws_t[:5], dws_f[:5], ["dW0", "db0", "dW1", "db1", "dW2"]):
        assert float((a - b_).norm()) <= 1e-3 * float(b_.norm()), name
        assert float((a - b_).abs().max()) <= 2e-3 * float(b_.abs().max()), name
    # ... and the kernels themselves are checked exactly: with the ReLU masks FIXED to the ones the forward stashed,
    # the gradient is a smooth function, and a float64 evaluation of it must agree to 1e-5 everywhere.
    attw, act0, act1 = [t.double().cpu() for t in stash]
    W0, b0, W1, b1, W2, b2 = [t.double().cpu() for t in cw]
    X, G = rows.double().cpu(), gup.double().cpu()
    h, t = X[:, :L], X[:, L]
    dw = (G.unsqueeze(1) * h).sum(-1) if pool else (G * h).sum(-1)                   # (B, L)
    ds = attw * (dw - (attw * dw).sum(1, keepdim=True))
    dz1 = ds.reshape(-1, 1) * W2 * (act1 > 0)
    dz0 = (dz1 @ W1) * (act0 > 0)
    Wab, Wt = W0[:, :D] + W0[:, D:2 * D], W0[:, 2 * D:] - W0[:, D:2 * D]
    direct = attw.unsqueeze(-1) * (G.unsqueeze(1) if pool else G)
    dtb = dz0.view(B, L, H1).sum(1)
    dWab, dWt = dz0.t() @ h.reshape(-1, D), dtb.t() @ t
    want = {"d_hist": (dz0 @ Wab).view(B, L, D) + direct, "d_tgt": dtb @ Wt, "dW0": torch.cat([dWab, dWab - dWt, dWt], 1),
            "db0": dtb.sum(0), "dW1": dz1.t() @ act0, "db1": dz1.sum(0), "dW2": ds.reshape(1, -1) @ act1}
    got = {"d_hist": d_t[:, :L], "d_tgt": d_t[:, L], "dW0": dws_t[0], "db0": dws_t[1], "dW1": dws_t[2], "db1": dws_t[3],
           "dW2": dws_t[4]}
    for name in want:
        ref = want[name]
        close(got[name].double(), ref, rtol=1e-5, atol=1e-5 * float(ref.abs().max()), msg=name + " (fixed masks, f64)")


# ------------------------------------------------------------------ tcgen05 3xTF32 A^T B (PNN "out")
@pytest.mark.parametrize("K,M,N", [(1000, 32, 32), (4096, 128, 256), (33, 8, 16), (70000, 16, 16), (257, 100, 72)])
def test_gemm_tn_3xtf32(K, M, N):
    ops = _ops()
    g = torch.Generator().manual_seed(K + M)
    A, B = torch.randn(K, M, generator=g), torch.randn(K, N, generator=g)
    want = (A.double().t() @ B.double())
    got = ops.gemm_tn(A.cuda(), B.cuda()).cpu().double()
    # 1e-5 relative to the scale of the sums (|a||b| mass), the bar a plain fp32 matmul meets as well
    tol = 1e-5 * float((A.abs().double().t() @ B.abs().double()).max())
    assert float((got - want).abs().max()) <= tol, (float((got - want).abs().max()), tol)
    # a single-pass TF32 product would miss this bar by two orders of magnitude (~1e-3 relative)


@pytest.mark.parametrize("M,K,N", [(1000, 64, 128), (128, 128, 64), (4097, 16, 128), (333, 128, 128), (50, 32, 1), (700, 64, 256)])
def test_gemm_nt_3xtf32_with_epilogue(M, K, N):
    ops = _ops()
    g = torch.Generator().manual_seed(M + K + N)
    A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    bias, group = torch.randn(N, generator=g), 7
    rowbias = torch.randn((M + group - 1) // group, N, generator=g)
    mask = torch.randn(M, N, generator=g)
    z = A.double() @ W.double().t() + bias.double() + rowbias.double().repeat_interleave(group, 0)[:M]
    want = torch.where(mask > 0, torch.relu(z), torch.zeros_like(z))
    got = ops.gemm_nt(A.cuda(), W.cuda(), bias=bias.cuda(), rowbias=rowbias.cuda(), rb_group=group, relu=True, mask=mask.cuda()).cpu().double()
    tol = 1e-5 * float((A.abs().double() @ W.abs().double().t()).max())
    assert float((got - want).abs().max()) <= tol
    plain = ops.gemm_nt(A.cuda(), W.cuda()).cpu().double()
    assert float((plain - A.double() @ W.double().t()).abs().max()) <= tol


def test_gemm_nt_rejects_weights_that_do_not_fit():
    ops = _ops()
    with pytest.raises(RuntimeError, match="shared memory"):
        ops.gemm_nt(torch.zeros(10, 192).cuda(), torch.zeros(128, 192).cuda())


def test_gemm_nt_grouped_rows():
    """A given as the first L rows of each (L+1)-row group of a (B, L+1, D) tensor -- the DIN history view."""
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    B, L, D, N = 37, 9, 64, 128
    rows = torch.randn(B, L + 1, D, generator=g).cuda()
    W = (torch.randn(N, D, generator=g) / 8).cuda()
    got = ops.gemm_nt(rows, W, M=B * L, a_rows=(L, (L + 1) * D, D))
    want = (rows[:, :L].reshape(B * L, D).double() @ W.double().t())
    assert float((got.double() - want).abs().max()) <= 1e-5 * float((rows.abs().max() * W.abs().sum(1).max()))


# ------------------------------------------------------------------ round-2 entry points
def test_sigmoid_bce_with_bias_and_gradient_sum():
    """rs_sigmoid_bce_bias: logit = cross + bias folded in, bias gradient = sum of d loss / d logit reduced beside the loss"""
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    for B in (1, 37, 5000, 300000):                   # 300000 > 1024 blocks x 256 threads: the grid-stride path
        cross = (torch.randn(B, generator=g) * 2).requires_grad_(True)
        bias = torch.tensor([0.37], requires_grad=True)
        y = (torch.rand(B, generator=g) < 0.3).float()
        loss = torch.nn.BCELoss()(torch.sigmoid(cross + bias), y)
        gc, gb = torch.autograd.grad(loss, (cross, bias))
        pred, l, gz, gsum = ops.sigmoid_bce(cross.detach().cuda(), y.cuda(), bias=bias.detach().cuda(), want_gsum=True)
        close(pred, torch.sigmoid(cross + bias).detach(), rtol=1e-6, atol=1e-7)
        close(l, loss.detach(), rtol=1e-5)
        close(gz, gc, rtol=1e-4, atol=1e-9)
        close(gsum, gb, rtol=1e-5, atol=1e-7)
        pred0, l0, gz0 = ops.sigmoid_bce((cross + bias).detach().cuda(), y.cuda())        # the plain entry point agrees
        assert torch.equal(pred0, pred) and torch.equal(gz0, gz) and torch.equal(l0, l)


@pytest.mark.parametrize("world,numel", [(1, 64), (2, 4 * 1000), (3, 4 * 333), (8, 4 * 4099)])
def test_replica_sgd_equals_allreduce_plus_sgd(world, numel):
    """rs_replica_sgd run by every rank in turn (the peers are just other buffers): every replica ends up with
    w - lr * (sum of the ranks' gradients), and the replicas are bit-identical."""
    ops = _ops()
    g = torch.Generator().manual_seed(world)
    w0 = torch.randn(numel, generator=g)
    grads = [torch.randn(numel, generator=g) for _ in range(world)]
    W = [w0.clone().cuda() for _ in range(world)]
    G = [x.cuda() for x in grads]
    for r in range(world):
        ops.replica_sgd([t.data_ptr() for t in W], [t.data_ptr() for t in G], numel, world, r, 0.25)
    acc = torch.zeros(numel, dtype=torch.float32)
    for x in grads:
        acc = acc + x                                   # rank order, fp32: the kernel's order
    want = w0 - 0.25 * acc
    for r in range(world):
        assert torch.equal(W[r], W[0])
        close(W[r], want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("mode", ["grad", "sgd"])
def test_streaming_update_half_sm_is_bit_identical(mode):
    """half_sm (one CTA per SM, room for a concurrent kernel) only changes the launch geometry"""
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    rows, n, W = 500, 20000, 416
    ids = (torch.rand(n, generator=g) ** 3 * rows).long().clamp_(max=rows - 1)
    G = torch.randn(n, W, generator=g).cuda()
    table = torch.randn(rows, W, generator=g)
    segs = ops.dedup_sort(ids.cuda(), 1, None, rows, max_width=W, reuse_workspace=False)
    outs = []
    for half in (False, True):
        if mode == "grad":
            out = torch.zeros(rows, W).cuda()
            ops.segment_update(segs, ops.RS_UPD_GRAD, W, 1, dense=G, dense_grad=out, half_sm=half)
        else:
            out = table.clone().cuda()
            ops.segment_update(segs, ops.RS_UPD_SGD, W, 1, dense=G, table=out, lr=0.1, half_sm=half)
        outs.append(out)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("F,D,B,mask", [(26, 16, 300, 0b10100000000100000000000100), (6, 8, 500, 0b100001), (40, 8, 100, (1 << 39) | (1 << 3) | 1)])
def test_ffm_split_stash(F, D, B, mask):
    """rs_ffm_fwd_peer with stash_split: the Jacobian rows of the masked fields go to their own tensor, field order kept"""
    ops = _ops()
    rows = [5 + 3 * f for f in range(F)]
    tabs, ids = make_fields(F, F * D, rows, B, seed=F + D)
    keep = [t.cuda() for t in tabs]
    T = ops.make_tables(keep)
    cross, full = ops.ffm_fwd(T, ids.cuda(), D, want_stash=True)
    peer = (1, 0, [keep[0].data_ptr()], 1)            # no direct fields: only the split is exercised
    c2, rest, split = ops.ffm_fwd(T, ids.cuda(), D, want_stash=True, peer=peer, split_mask=mask)
    big = [f for f in range(F) if (mask >> f) & 1]
    small = [f for f in range(F) if not (mask >> f) & 1]
    assert torch.equal(c2, cross)
    assert torch.equal(split, full[:, big]) and torch.equal(rest, full[:, small])
