"""Edge cases the reference tolerates: empty batches, single samples, single fields, out-of-range ids, max widths."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from deeplearningrecommendationsystem_b200 import ops
    return ops


def test_empty_batches_are_no_ops():
    ops = _ops()
    tab = torch.randn(10, 16).cuda()
    T = ops.make_tables([tab, tab])
    ids = torch.zeros(0, 2, dtype=torch.int64).cuda()
    assert ops.gather_rows(T, ids).shape == (0, 2, 16)
    out = ops.fields_fwd(T, 0, "cuda", ids=ids, cross=True, concat=True, stash=True, dot2=True)
    assert out["cross"].shape == (0,) and out["concat"].shape == (0, 32)
    E = torch.zeros(0, 2, 16).cuda()
    assert ops.fields_bwd(ops.dummy_tables(2, 16), 0, "cuda", dense_in=E, g_cross=torch.zeros(0).cuda()).shape == (0, 2, 16)
    cross, stash = ops.ffm_fwd(ops.make_tables([torch.randn(5, 32).cuda()] * 2), ids, 16)
    assert cross.shape == (0,) and stash.shape == (0, 2, 32)
    pooled, attw = ops.afm_fwd(torch.zeros(0, 3, 8).cuda(), torch.zeros(8, 4).cuda(), torch.zeros(4).cuda(), torch.zeros(4, 1).cuda())
    assert pooled.shape == (0, 8)
    ops.check_status()


def test_empty_batch_through_a_model():
    from deeplearningrecommendationsystem_b200.model import DeepFM, MatrixFactorization
    m = DeepFM(10, 12, [16, 8, 1], 8).cuda()
    assert m(torch.zeros(0, 45).cuda()).shape == (0, 1)
    mf = MatrixFactorization(10, 12, 8).cuda()
    assert mf(torch.zeros(0, dtype=torch.int64).cuda(), torch.zeros(0, dtype=torch.int64).cuda()).shape == (0,)


def test_single_sample_and_repeated_id():
    """B = 1, and a batch where every lookup hits the same row (one segment of B lookups, many chunks)."""
    from deeplearningrecommendationsystem_b200.model import MatrixFactorization
    from oracle import ml100k, interactions as OI
    torch.manual_seed(0)
    m = MatrixFactorization(7, 9, 16)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda()
    for B in (1, 5000):
        u = torch.full((B,), 3, dtype=torch.int64)
        i = torch.full((B,), 4, dtype=torch.int64)
        y = (torch.arange(B) % 2).float()
        m.zero_grad()
        p = m(u.cuda(), i.cuda())
        torch.nn.BCELoss()(p, y.cuda()).backward()
        pred, loss, grads = ml100k.loss_and_grads("mf", sd, [u, i], y)
        np.testing.assert_allclose(p.detach().cpu().numpy(), pred.numpy(), rtol=1e-5, atol=1e-6)
        for k, v in m.named_parameters():
            np.testing.assert_allclose(v.grad.cpu().numpy(), grads[k].numpy(), rtol=1e-5, atol=1e-6, err_msg=f"{k} B={B}")


def test_out_of_range_id_raises_index_error_like_the_reference():
    from deeplearningrecommendationsystem_b200 import ops
    from deeplearningrecommendationsystem_b200.model import MatrixFactorization
    m = MatrixFactorization(5, 5, 8).cuda()
    m(torch.tensor([1, 7]).cuda(), torch.tensor([0, 1]).cuda())
    with pytest.raises(IndexError):
        ops.check_status()
    m(torch.tensor([1, -1]).cuda(), torch.tensor([0, 1]).cuda())
    with pytest.raises(IndexError):
        ops.check_status()


def test_trainer_turns_a_bad_id_into_index_error():
    """the product path polls the device status word itself: valid/test passes and metrics() always, train_loop every
    RS_CHECK_EVERY steps (nothing but tests used to read the flag)"""
    from deeplearningrecommendationsystem_b200.model import MatrixFactorization
    from deeplearningrecommendationsystem_b200.trainer import Trainer, trainer as tmod
    m = MatrixFactorization(5, 5, 8).cuda()
    tr = Trainer(m, torch.nn.BCELoss(), torch.optim.SGD(m.parameters(), lr=0.1))
    y = torch.tensor([0.0, 1.0]).cuda()
    good, bad = (torch.tensor([1, 2]).cuda(), torch.tensor([0, 1]).cuda()), (torch.tensor([1, 9]).cuda(), torch.tensor([0, 1]).cuda())
    tr.valid_loop(*good, valid_rating=y)
    with pytest.raises(IndexError):
        tr.valid_loop(*bad, valid_rating=y)
    with pytest.raises(IndexError):
        tr.test_loop(*bad, test_rating=y)
    old = tmod.CHECK_EVERY
    tmod.CHECK_EVERY = 2
    try:
        tr._steps = 0
        tr.train_loop(*bad, train_rating=y)          # step 1: not polled yet
        with pytest.raises(IndexError):
            tr.train_loop(*good, train_rating=y)     # step 2: the flag raised by step 1 surfaces
    finally:
        tmod.CHECK_EVERY = old
    tr.valid_loop(*good, valid_rating=y)             # the flag was cleared by the raise


def test_graphed_adam_is_refused_not_silently_frozen():
    """Adam's bias correction comes from a host step counter, so a captured Adam step would replay with
    it frozen; GraphedTrainStep refuses, SGD captures and N graphed steps == N eager steps."""
    from deeplearningrecommendationsystem_b200.graph import GraphedTrainStep
    from deeplearningrecommendationsystem_b200.nfield import FieldFM
    from deeplearningrecommendationsystem_b200.optim import DenseAdam, FusedRowOptimizer
    cards = [50, 7, 300]
    g = torch.Generator().manual_seed(0)
    ids = torch.stack([torch.randint(0, c, (256,), generator=g) for c in cards], dim=1).cuda()
    y = (torch.rand(256, 1, generator=g) < 0.3).float().cuda()
    m = FieldFM(cards, 16, seed=1, device="cuda")
    for opt in (FusedRowOptimizer(m, None, lr=0.1, kind="adam"), FusedRowOptimizer(m, DenseAdam([m.bias]), lr=0.1),
                FusedRowOptimizer(m, torch.optim.Adam([m.bias]), lr=0.1)):
        gs = GraphedTrainStep(m, torch.nn.BCELoss(), opt, warmup=0)
        with pytest.raises(RuntimeError, match="captur"):
            gs(ids, rating=y)
    a, b = FieldFM(cards, 16, seed=2, device="cuda"), FieldFM(cards, 16, seed=2, device="cuda")
    oa = FusedRowOptimizer(a, torch.optim.SGD([a.bias], lr=0.1), lr=0.1)
    ob = FusedRowOptimizer(b, torch.optim.SGD([b.bias], lr=0.1), lr=0.1)
    gs = GraphedTrainStep(a, torch.nn.BCELoss(), oa, warmup=1)
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    tr = Trainer(b, torch.nn.BCELoss(), ob)
    for _ in range(5):
        gs(ids, rating=y)
        tr.train_loop(ids, train_rating=y)
    torch.cuda.synchronize()
    assert torch.equal(a.weight.data, b.weight.data) and torch.equal(a.bias.data, b.bias.data)


def test_unsupported_shapes_fail_loudly():
    ops = _ops()
    with pytest.raises(RuntimeError, match="power of two"):
        ops.fields_fwd(ops.dummy_tables(3, 12), 4, "cuda", dense_in=torch.zeros(4, 3, 12).cuda(), cross=True)
    with pytest.raises(TypeError):
        ops.gather_rows(ops.make_tables([torch.zeros(4, 8).cuda()]), torch.zeros(3, 1, dtype=torch.int32).cuda())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gather_rows(ops.make_tables([torch.zeros(4, 8).cuda()]), torch.zeros(3, 1, dtype=torch.int64))


def test_widest_supported_rows():
    """W = 2048 floats (the segment kernels' limit) and D = 256 (the field kernels' limit)."""
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 30, (200,), generator=g)
    G = torch.randn(200, 2048, generator=g)
    segs = ops.dedup_sort(ids.cuda(), 1, None, 30, max_width=2048, reuse_workspace=False)
    dense = torch.zeros(30, 2048).cuda()
    ops.segment_update(segs, ops.RS_UPD_GRAD, 2048, 1, dense=G.cuda(), dense_grad=dense)
    np.testing.assert_allclose(dense.cpu().numpy(), torch.zeros(30, 2048).index_add_(0, ids, G).numpy(), rtol=1e-5, atol=1e-5)
    E = torch.randn(9, 3, 256, generator=g)
    out = ops.fields_fwd(ops.dummy_tables(3, 256), 9, "cuda", dense_in=E.cuda(), cross=True, bi=True)
    s = E.sum(1)
    np.testing.assert_allclose(out["bi"].cpu().numpy(), (0.5 * (s * s - (E * E).sum(1))).numpy(), rtol=1e-5, atol=1e-4)
