"""N-field FM / FFM train steps (the BASELINE.json configs[1] path) against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import nfield as onf

pytestmark = pytest.mark.gpu

CARDS = [3, 50, 7, 1000, 24, 12, 301, 5]


def _setup(kind, D, B, seed, zipf):
    from deeplearningrecommendationsystem_b200.nfield import FieldFFM, FieldFM
    g = torch.Generator().manual_seed(seed)
    model = (FieldFM if kind == "fm" else FieldFFM)(CARDS, D, fused=True, seed=seed, device="cuda")
    if zipf:
        ids = torch.stack([(torch.rand(B, generator=g) ** 3 * c).long().clamp_(max=c - 1) for c in CARDS], dim=1)
    else:
        ids = torch.stack([torch.randint(0, c, (B,), generator=g) for c in CARDS], dim=1)
    y = (torch.rand(B, 1, generator=g) < 0.3).float()
    return model, ids, y


@pytest.mark.parametrize("kind,D", [("fm", 16), ("ffm", 4), ("ffm", 16), ("fm", 64)])
@pytest.mark.parametrize("zipf", [False, True])
def test_fused_sgd_steps_match_oracle(kind, D, zipf):
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    B, lr = 700, 0.5
    model, ids, y = _setup(kind, D, B, 21, zipf)
    table = model.weight.detach().cpu().clone()
    bias = model.bias.detach().cpu().clone()
    offsets = torch.tensor(model.offsets_host)
    opt = FusedRowOptimizer(model, torch.optim.SGD([model.bias], lr=lr), lr=lr, kind="sgd")
    tr = Trainer(model, torch.nn.BCELoss(), opt)
    for step in range(3):
        tr.train_loop(ids.cuda(), train_rating=y.cuda())
        pred, loss = onf.train_step(kind, table, bias, ids, offsets, y[:, 0], lr)
        np.testing.assert_allclose(tr.predictions_train.detach().cpu().numpy()[:, 0], pred.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(tr.train_loss.item(), loss.item(), rtol=1e-5)
        np.testing.assert_allclose(model.weight.detach().cpu().numpy(), table.numpy(), rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(model.bias.detach().cpu().numpy(), bias.numpy(), rtol=1e-5, atol=1e-6)
    from deeplearningrecommendationsystem_b200 import ops
    ops.check_status()


C2_CARDS = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992, 5461306, 10, 5652,
            2173, 4, 7046547, 18, 15, 286181, 105, 142572]


@pytest.mark.parametrize("kind", ["fm", "ffm"])
@pytest.mark.parametrize("zipf", [False, True])
def test_c2_shape_train_steps_match_oracle(kind, zipf):
    """BASELINE.json configs[1] at its own shape -- F = 26 Criteo fields, D = 16 (FFM rows of 416 floats, the TMA-fed
    ffm_fwd_kernel and seg_stream_kernel) -- with every cardinality capped at 3000 rows so the CPU oracle's dense
    gradient stays small; three fused SGD steps through Trainer.train_loop vs oracle/nfield.train_step."""
    from deeplearningrecommendationsystem_b200 import ops
    from deeplearningrecommendationsystem_b200.nfield import FieldFFM, FieldFM
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    cards = [min(c, 3000) for c in C2_CARDS]
    B, D, lr = 2048, 16, 0.5
    g = torch.Generator().manual_seed(26)
    model = (FieldFM if kind == "fm" else FieldFFM)(cards, D, fused=True, seed=7, device="cuda")
    if zipf:
        ids = torch.stack([(torch.rand(B, generator=g) ** 3 * c).long().clamp_(max=c - 1) for c in cards], dim=1)
    else:
        ids = torch.stack([torch.randint(0, c, (B,), generator=g) for c in cards], dim=1)
    y = (torch.rand(B, 1, generator=g) < 0.3).float()
    table, bias = model.weight.detach().cpu().clone(), model.bias.detach().cpu().clone()
    offsets = torch.tensor(model.offsets_host)
    opt = FusedRowOptimizer(model, torch.optim.SGD([model.bias], lr=lr), lr=lr, kind="sgd")
    tr = Trainer(model, torch.nn.BCELoss(), opt)
    for step in range(3):
        tr.train_loop(ids.cuda(), train_rating=y.cuda())
        pred, loss = onf.train_step(kind, table, bias, ids, offsets, y[:, 0], lr)
        np.testing.assert_allclose(tr.predictions_train.detach().cpu().numpy()[:, 0], pred.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(tr.train_loss.item(), loss.item(), rtol=1e-5)
        np.testing.assert_allclose(model.weight.detach().cpu().numpy(), table.numpy(), rtol=1e-5, atol=2e-6)
    ops.check_status()


@pytest.mark.parametrize("kind,D", [("fm", 16), ("ffm", 8)])
def test_dense_grad_mode_matches_autograd(kind, D):
    """fused=False: the table is an ordinary dense-gradient Parameter (reference semantics, any torch optimizer)."""
    from deeplearningrecommendationsystem_b200.nfield import FieldFFM, FieldFM
    g = torch.Generator().manual_seed(5)
    B = 300
    model = (FieldFM if kind == "fm" else FieldFFM)(CARDS, D, fused=False, seed=3, device="cuda")
    ids = torch.stack([torch.randint(0, c, (B,), generator=g) for c in CARDS], dim=1)
    y = (torch.rand(B, 1, generator=g) < 0.3).float()
    loss = torch.nn.BCELoss()(model(ids.cuda()), y.cuda())
    loss.backward()
    table = model.weight.detach().cpu().clone().requires_grad_(True)
    bias = model.bias.detach().cpu().clone().requires_grad_(True)
    offsets = torch.tensor(model.offsets_host)
    logit = (onf.fm_logit if kind == "fm" else onf.ffm_logit)(table, ids, offsets, bias)
    want = torch.nn.BCELoss()(torch.sigmoid(logit), y[:, 0])
    want.backward()
    np.testing.assert_allclose(loss.item(), want.item(), rtol=1e-5)
    np.testing.assert_allclose(model.weight.grad.cpu().numpy(), table.grad.numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(model.bias.grad.cpu().numpy(), bias.grad.numpy(), rtol=1e-5, atol=1e-7)


def test_inference_path_needs_no_stash():
    model, ids, y = _setup("ffm", 8, 100, 2, False)
    with torch.no_grad():
        p1 = model(ids.cuda())
    p2 = model(ids.cuda())
    assert torch.equal(p1, p2.detach()) and p1.shape == (100, 1)


def test_fused_adam_matches_row_oracle():
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from oracle import optim as oo, interactions as OI
    model, ids, y = _setup("fm", 16, 500, 8, True)
    table = model.weight.detach().cpu().clone()
    m, v = torch.zeros_like(table), torch.zeros_like(table)
    offsets = torch.tensor(model.offsets_host)
    opt = FusedRowOptimizer(model, None, lr=1e-2, kind="adam")
    for step in (1, 2, 3):
        opt.zero_grad()
        loss = torch.nn.BCELoss()(model(ids.cuda()), y.cuda())
        loss.backward()
        opt.step()
        E = onf.gather_fields(table, ids, offsets).requires_grad_(True)
        l = OI.bce(torch.sigmoid(OI.fm_second_order(E) + model.bias.detach().cpu()), y[:, 0])
        (G,) = torch.autograd.grad(l, E)
        oo.adam_rows(table, m, v, (ids + offsets).reshape(-1), G.reshape(-1, 16), step, lr=1e-2)
        np.testing.assert_allclose(model.weight.detach().cpu().numpy(), table.numpy(), rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("kind", ["sgd", "adam"])
def test_field_mf_matches_mf_oracle(kind):
    """FieldMF (one concatenated table, fused row update) == the reference MF arithmetic (oracle.ml100k.mf)."""
    from deeplearningrecommendationsystem_b200.nfield import FieldMF
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from oracle import ml100k, optim as oo, interactions as OI
    nu, ni, D, B = 300, 500, 64, 2000
    g = torch.Generator().manual_seed(3)
    u = (torch.rand(B, generator=g) ** 2 * nu).long()
    i = torch.randint(0, ni, (B,), generator=g)
    y = (torch.rand(B, generator=g) < 0.3).float()
    m = FieldMF(nu, ni, D, fused=True, seed=5, device="cuda")
    U, V = m.weight.detach().cpu()[:nu].clone(), m.weight.detach().cpu()[nu:].clone()
    opt = FusedRowOptimizer(m, None, lr=0.05, kind=kind)
    state = {k: (torch.zeros_like(t), torch.zeros_like(t)) for k, t in (("U", U), ("V", V))}
    for step in (1, 2, 3):
        opt.zero_grad()
        pred = m(u.cuda(), i.cuda())
        assert pred.shape == (B,)
        loss = torch.nn.BCELoss()(pred, y.cuda())
        loss.backward()
        opt.step()
        sd = {"user_embeddings.weight": U.clone().requires_grad_(True), "item_embeddings.weight": V.clone().requires_grad_(True)}
        p = ml100k.mf(sd, u, i)
        l = OI.bce(p, y)
        gU, gV = torch.autograd.grad(l, list(sd.values()))
        np.testing.assert_allclose(pred.detach().cpu().numpy(), p.detach().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(loss.item(), l.item(), rtol=1e-5)
        if kind == "sgd":
            U -= 0.05 * gU
            V -= 0.05 * gV
        else:
            for name, tab, gr, ids in (("U", U, gU, u), ("V", V, gV, i)):
                touched = torch.unique(ids)
                oo.adam_rows(tab, state[name][0], state[name][1], touched, gr[touched], step, lr=0.05)
        got = m.weight.detach().cpu()
        np.testing.assert_allclose(got[:nu].numpy(), U.numpy(), rtol=2e-5, atol=2e-4 if kind == "adam" else 2e-6)
        np.testing.assert_allclose(got[nu:].numpy(), V.numpy(), rtol=2e-5, atol=2e-4 if kind == "adam" else 2e-6)


@pytest.mark.parametrize("kind", ["pnn", "afm"])
def test_field_pnn_afm_train_steps(kind):
    """N-field PNN (inner product) / AFM: fused sparse SGD on the tables + torch SGD on the dense layers == the CPU
    restatement with dense embedding gradients (oracle.interactions) -- predictions, loss and every parameter."""
    from deeplearningrecommendationsystem_b200.nfield import FieldAFM, FieldPNN
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    from oracle import interactions as OI
    cards, D, B, lr = [7, 300, 40, 1000, 5, 64, 12], 16, 600, 0.2
    g = torch.Generator().manual_seed(4)
    ids = torch.stack([(torch.rand(B, generator=g) ** 2 * c).long().clamp_(max=c - 1) for c in cards], dim=1)
    y = (torch.rand(B, 1, generator=g) < 0.4).float()
    torch.manual_seed(0)
    m = (FieldPNN(cards, D, [32, 16, 8], seed=2, device="cuda") if kind == "pnn" else FieldAFM(cards, D, 8, seed=2, device="cuda"))
    dense = [p for n, p in m.named_parameters() if p.requires_grad]
    ref = {n: p.detach().cpu().clone().requires_grad_(True) for n, p in m.named_parameters() if n != "bias"}
    offs = torch.tensor(m.offsets_host)
    tr = Trainer(m, torch.nn.BCELoss(), FusedRowOptimizer(m, torch.optim.SGD(dense, lr=lr), lr=lr))
    for _ in range(3):
        tr.train_loop(ids.cuda(), train_rating=y.cuda())
        E = ref["weight"][ids + offs]
        if kind == "pnn":
            h = E.flatten(1) @ ref["linear1.weight"].t() + ref["linear1.bias"] + OI.inner_products(E) @ ref["linear2.weight"].t() + ref["linear2.bias"]
            k = 0
            while f"dnn_network.{k}.weight" in ref:
                h = torch.relu(h @ ref[f"dnn_network.{k}.weight"].t() + ref[f"dnn_network.{k}.bias"])
                k += 1
            pred = torch.sigmoid(h @ ref["output.weight"].t() + ref["output.bias"])
        else:
            pooled = OI.afm_pool(E, ref["attention_W"], ref["attention_b"], ref["attention_h"])
            pred = torch.sigmoid(pooled @ ref["output_layer.weight"].t() + ref["output_layer.bias"])
        loss = OI.bce(pred, y)
        grads = torch.autograd.grad(loss, list(ref.values()))
        np.testing.assert_allclose(tr.predictions_train.detach().cpu().numpy(), pred.detach().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(tr.train_loss.item(), loss.item(), rtol=1e-5)
        with torch.no_grad():
            for (n, p), gr in zip(ref.items(), grads):
                p -= lr * gr
        for n, p in m.named_parameters():
            if n != "bias":
                np.testing.assert_allclose(p.detach().cpu().numpy(), ref[n].detach().numpy(), rtol=1e-5, atol=2e-6, err_msg=n)


@pytest.mark.parametrize("kind", ["fm", "ffm", "mf"])
def test_graphed_step_equals_eager(kind):
    """CUDA-graph replay of the whole train step gives bit-identical tables to the eager Trainer loop."""
    from deeplearningrecommendationsystem_b200.graph import GraphedTrainStep
    from deeplearningrecommendationsystem_b200.nfield import FieldFFM, FieldFM, FieldMF
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    B, lr = 512, 0.3

    def build():
        if kind == "mf":
            return FieldMF(200, 300, 16, seed=1, device="cuda")
        return (FieldFM if kind == "fm" else FieldFFM)(CARDS, 8, seed=1, device="cuda")

    def batch(k):
        g = torch.Generator().manual_seed(40 + k)
        if kind == "mf":
            return (torch.randint(0, 200, (B,), generator=g).cuda(), torch.randint(0, 300, (B,), generator=g).cuda()), \
                (torch.rand(B, generator=g) < 0.3).float().cuda()
        ids = torch.stack([torch.randint(0, c, (B,), generator=g) for c in CARDS], dim=1).cuda()
        return (ids,), (torch.rand(B, 1, generator=g) < 0.3).float().cuda()

    def opt_for(m):
        dense = torch.optim.SGD([m.bias], lr=lr) if m.bias.requires_grad else None
        return FusedRowOptimizer(m, dense, lr=lr)

    m1, m2 = build(), build()
    tr = Trainer(m1, torch.nn.BCELoss(), opt_for(m1))
    gs = GraphedTrainStep(m2, torch.nn.BCELoss(), opt_for(m2), warmup=1)
    for k in range(5):
        ins, y = batch(k)
        tr.train_loop(*ins, train_rating=y)
        pred, loss = gs(*ins, rating=y)
        assert torch.equal(pred, tr.predictions_train) and torch.equal(loss, tr.train_loss)
    assert torch.equal(m1.weight, m2.weight) and torch.equal(m1.bias, m2.bias)


@pytest.mark.parametrize("cards,D,B,wd", [
    ([min(c, 3000) for c in C2_CARDS], 16, 4096, 0.0),          # the C2 row shape (416 floats), heavy duplicates + single rows
    ([min(c, 200000) for c in C2_CARDS], 16, 3000, 1e-3),       # mostly rows looked up once, weight decay
    ([3, 50, 7, 1000, 24], 4, 700, 0.0),                        # 20-float rows
    ([5 + 37 * k for k in range(40)], 8, 1500, 0.0),            # F = 40 > 32 fields (two id registers per lane)
    ([2, 3], 64, 10000, 0.0),                                   # two tiny tables: every row spans many chunks
    ([70000 if k % 7 == 0 else 5 + 37 * k for k in range(40)], 8, 1200, 0.0),   # F > 32 with six cold (>= 2^16 rows) fields
    ([70000, 90000, 66000, 131072, 80000], 8, 3000, 1e-3),      # every field cold: all slices come from the cold-slice stash
    ([65536, 9, 100000], 32, 2500, 0.0),                        # D = 32
])
def test_ffm_stash_free_step_is_bit_identical_to_stash_step(cards, D, B, wd, monkeypatch):
    """rs_ffm_bwd_update (gradient rows recomputed from the table: no Jacobian stash) against rs_ffm_fwd(stash) +
    rs_segment_update: same per-element arithmetic, chunking and combine order => the same bits after every step."""
    from deeplearningrecommendationsystem_b200 import ops
    from deeplearningrecommendationsystem_b200.nfield import FieldFFM
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    g = torch.Generator().manual_seed(3)
    ids = torch.stack([torch.randint(0, c, (B,), generator=g) for c in cards], dim=1).cuda()
    y = (torch.rand(B, 1, generator=g) < 0.3).float().cuda()
    lr = 0.3
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("RS_FFM_RECOMPUTE", mode)
        m = FieldFFM(cards, D, fused=True, seed=11, device="cuda")
        tr = Trainer(m, torch.nn.BCELoss(), FusedRowOptimizer(m, torch.optim.SGD([m.bias], lr=lr), lr=lr, weight_decay=wd))
        assert m._recompute == (mode == "1")
        launches = ops.launches()
        for _ in range(3):
            tr.train_loop(ids, train_rating=y)
        out[mode] = (m.weight.detach().clone(), tr.predictions_train.detach().clone(), ops.launches() - launches)
    ops.check_status()
    assert torch.equal(out["1"][1], out["0"][1])
    assert torch.equal(out["1"][0], out["0"][0])
    assert out["1"][2] != out["0"][2]           # the two runs really took different kernels


def test_ffm_second_forward_before_step_keeps_its_stash():
    """two forwards, one step: the second record must not be recomputed from a table the first record already updated"""
    from deeplearningrecommendationsystem_b200.nfield import FieldFFM
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    g = torch.Generator().manual_seed(4)
    B = 500
    ids = [torch.stack([torch.randint(0, c, (B,), generator=g) for c in CARDS], dim=1).cuda() for _ in range(2)]
    y = (torch.rand(B, 1, generator=g) < 0.3).float().cuda()
    res = []
    for rec in ("1", "0"):
        import os
        os.environ["RS_FFM_RECOMPUTE"] = rec
        try:
            m = FieldFFM(CARDS, 8, fused=True, seed=2, device="cuda")
            opt = FusedRowOptimizer(m, torch.optim.SGD([m.bias], lr=0.2), lr=0.2)
        finally:
            os.environ.pop("RS_FFM_RECOMPUTE")
        opt.zero_grad()
        loss = torch.nn.BCELoss()(m(ids[0]), y) + torch.nn.BCELoss()(m(ids[1]), y)
        loss.backward()
        opt.step()
        res.append(m.weight.detach().clone())
    assert torch.equal(res[0], res[1])


@pytest.mark.parametrize("kind", ["fm", "ffm", "mf"])
def test_trainer_fused_sigmoid_bce_matches_generic_path(kind, monkeypatch):
    """Trainer.train_loop with nn.BCELoss on a model that exposes train_logit runs sigmoid + loss + both backward passes as
    rs_sigmoid_bce; the generic autograd path (RS_FUSED_BCE=0) must give the same predictions, loss and updated rows."""
    from deeplearningrecommendationsystem_b200.nfield import FieldFFM, FieldFM, FieldMF
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    g = torch.Generator().manual_seed(12)
    B = 900
    out = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("RS_FUSED_BCE", fused)
        if kind == "mf":
            m = FieldMF(300, 500, 16, seed=4, device="cuda")
            ins = (torch.randint(0, 300, (B,), generator=torch.Generator().manual_seed(1)).cuda(),
                   torch.randint(0, 500, (B,), generator=torch.Generator().manual_seed(2)).cuda())
            y = (torch.rand(B, generator=torch.Generator().manual_seed(3)) < 0.3).float().cuda()
            opt = FusedRowOptimizer(m, None, lr=0.3)
        else:
            m = (FieldFM if kind == "fm" else FieldFFM)(CARDS, 8, seed=4, device="cuda")
            ins = (torch.stack([torch.randint(0, c, (B,), generator=torch.Generator().manual_seed(c)) for c in CARDS], dim=1).cuda(),)
            y = (torch.rand(B, 1, generator=torch.Generator().manual_seed(3)) < 0.3).float().cuda()
            opt = FusedRowOptimizer(m, torch.optim.SGD([m.bias], lr=0.3), lr=0.3)
        tr = Trainer(m, torch.nn.BCELoss(), opt)
        for _ in range(2):
            tr.train_loop(*ins, train_rating=y)
        assert tr.predictions_train.shape == y.shape
        out[fused] = (tr.predictions_train.detach().cpu().numpy(), float(tr.train_loss), m.weight.detach().cpu().numpy(),
                      m.bias.detach().cpu().numpy())
    for a, b in zip(out["1"], out["0"]):
        np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-7)
