"""Opt-in fused sparse rows for the id-indexed drop-in modules (DIN, DIEN, MatrixFactorization, NeuralCF), and the
large-table NeuralCF of BASELINE.json configs[4] (nfield.FieldNeuralCF), single-GPU and row-sharded.

Reference semantics = nn.Embedding dense gradient + dense optimizer sweep (model/din.py:35-36, trainer/trainer.py:38-39).
With plain SGD the fused mode must give the SAME parameters (untouched rows have zero gradient); the check is against the
same module in its default dense-gradient mode, which the golden tests pin to the reference."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(name, B, g):
    if name in ("din", "dien"):
        return (torch.randint(0, 500, (B, 12), generator=g).cuda(), torch.randint(0, 500, (B,), generator=g).cuda())
    return (torch.randint(0, 300, (B,), generator=g).cuda(), torch.randint(0, 500, (B,), generator=g).cuda())


def _build(name):
    from deeplearningrecommendationsystem_b200 import model as M
    torch.manual_seed(3)
    if name == "din":
        return M.DIN(500, 16)
    if name == "dien":
        return M.DIEN(500, 16)
    if name == "mf":
        return M.MatrixFactorization(300, 500, 32)
    return M.NeuralCF(300, 500, 16, [32, 16, 8])


@pytest.mark.parametrize("name", ["din", "dien", "mf", "neuralcf"])
def test_fused_rows_equal_dense_sgd(name):
    from deeplearningrecommendationsystem_b200 import ops
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    lr, B = 0.2, 400
    dense = _build(name).cuda()
    fused = copy.deepcopy(dense).fuse_embedding_updates()
    keys = list(dense.state_dict())
    assert list(fused.state_dict()) == keys                              # same state_dict contract
    assert all(fused.state_dict()[k].shape == dense.state_dict()[k].shape for k in keys)
    rest = [p for p in fused.parameters() if p.requires_grad]
    assert len(rest) < len(list(fused.parameters()))                     # the tables left autograd (MF has nothing else)
    td = Trainer(dense, torch.nn.BCELoss(), torch.optim.SGD(dense.parameters(), lr=lr))
    tf = Trainer(fused, torch.nn.BCELoss(), FusedRowOptimizer(fused, torch.optim.SGD(rest, lr=lr) if rest else None, lr=lr))
    g = torch.Generator().manual_seed(11)
    for _ in range(3):
        ins = _inputs(name, B, g)
        y = (torch.rand(B, generator=g) < 0.4).float().cuda()
        if name != "mf":
            y = y.view(-1, 1)
        td.train_loop(*ins, train_rating=y)
        tf.train_loop(*ins, train_rating=y)
        np.testing.assert_allclose(tf.predictions_train.detach().cpu().numpy(), td.predictions_train.detach().cpu().numpy(),
                                   rtol=1e-5, atol=1e-6)
    for k in keys:
        np.testing.assert_allclose(fused.state_dict()[k].cpu().numpy(), dense.state_dict()[k].cpu().numpy(), rtol=1e-5, atol=2e-6,
                                   err_msg=k)
    for p in fused.parameters():
        assert p.requires_grad or p.grad is None                          # nothing table-sized was ever allocated
    with torch.no_grad():                                                 # inference path of the fused module
        np.testing.assert_allclose(fused(*ins).cpu().numpy(), dense(*ins).cpu().numpy(), rtol=1e-5, atol=1e-6)
    ops.check_status()


def test_fused_rows_lazy_adam_matches_row_oracle():
    """kind='adam' on a fused DIN: touched rows follow torch.optim.Adam's arithmetic (oracle/optim.adam_rows, 'parity
    unpinned' by construction for untouched rows -- see optim.py)"""
    from deeplearningrecommendationsystem_b200 import model as M
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from oracle import optim as oo
    torch.manual_seed(0)
    m = M.DIN(200, 16).cuda().fuse_embedding_updates()
    table0 = m.item_embedding.weight.detach().cpu().clone()
    rest = [p for p in m.parameters() if p.requires_grad]
    opt = FusedRowOptimizer(m, torch.optim.SGD(rest, lr=0.0), lr=0.01, kind="adam")
    g = torch.Generator().manual_seed(1)
    hist, tgt = torch.randint(0, 200, (64, 5), generator=g), torch.randint(0, 200, (64,), generator=g)
    y = (torch.rand(64, 1, generator=g) < 0.5).float()
    # oracle gradient of the same forward through the dense-mode module
    ref = M.DIN(200, 16).cuda()
    ref.load_state_dict(m.state_dict())
    torch.nn.BCELoss()(ref(hist.cuda(), tgt.cuda()), y.cuda()).backward()
    gdense = ref.item_embedding.weight.grad.cpu()
    opt.zero_grad()
    torch.nn.BCELoss()(m(hist.cuda(), tgt.cuda()), y.cuda()).backward()
    opt.step()
    ids = torch.cat([hist, tgt.unsqueeze(1)], dim=1).reshape(-1)
    touched = torch.unique(ids)
    want = table0.clone()
    mm, vv = torch.zeros_like(want), torch.zeros_like(want)
    oo.adam_rows(want, mm, vv, touched, gdense[touched], 1, lr=0.01)
    np.testing.assert_allclose(m.item_embedding.weight.detach().cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)


def _ncf_pair(device="cuda"):
    """FieldNeuralCF and the drop-in NeuralCF with identical weights"""
    from deeplearningrecommendationsystem_b200 import model as M
    from deeplearningrecommendationsystem_b200.nfield import FieldNeuralCF
    torch.manual_seed(5)
    ref = M.NeuralCF(300, 500, 16, [32, 16, 8]).to(device)
    f = FieldNeuralCF(300, 500, 16, [32, 16, 8], seed=1, device=device)
    with torch.no_grad():
        f.gmf.weight.copy_(torch.cat([ref.GMF_Embedding_User.weight, ref.GMF_Embedding_Item.weight]))
        f.mlp.weight.copy_(torch.cat([ref.MLP_Embedding_User.weight, ref.MLP_Embedding_Item.weight]))
        for a, b in zip(f.dnn_network, ref.dnn_network):
            a.load_state_dict(b.state_dict())
        f.linear.load_state_dict(ref.linear.state_dict())
        f.linear2.load_state_dict(ref.linear2.state_dict())
    return f, ref


def test_field_neuralcf_equals_dropin_neuralcf():
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    lr, B = 0.2, 500
    f, ref = _ncf_pair()
    dense = [p for p in f.parameters() if p.requires_grad]
    tf = Trainer(f, torch.nn.BCELoss(), FusedRowOptimizer(f, torch.optim.SGD(dense, lr=lr), lr=lr))
    tr = Trainer(ref, torch.nn.BCELoss(), torch.optim.SGD(ref.parameters(), lr=lr))
    g = torch.Generator().manual_seed(2)
    for _ in range(3):
        u, i = torch.randint(0, 300, (B,), generator=g).cuda(), torch.randint(0, 500, (B,), generator=g).cuda()
        y = (torch.rand(B, 1, generator=g) < 0.4).float().cuda()
        tf.train_loop(u, i, train_rating=y)
        tr.train_loop(u, i, train_rating=y)
        np.testing.assert_allclose(tf.predictions_train.detach().cpu().numpy(), tr.predictions_train.detach().cpu().numpy(),
                                   rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(f.gmf.weight.cpu().numpy(),
                               torch.cat([ref.GMF_Embedding_User.weight, ref.GMF_Embedding_Item.weight]).detach().cpu().numpy(),
                               rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(f.mlp.weight.cpu().numpy(),
                               torch.cat([ref.MLP_Embedding_User.weight, ref.MLP_Embedding_Item.weight]).detach().cpu().numpy(),
                               rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(f.linear2.weight.detach().cpu().numpy(), ref.linear2.weight.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_field_neuralcf_equals_single_gpu(world):
    """configs[4]: NeuralCF with both table pairs row-sharded (virtual ranks, one shared exchange plan per batch) ==
    the unsharded module on the concatenated global batch"""
    from deeplearningrecommendationsystem_b200 import dist as rsdist, ops
    from deeplearningrecommendationsystem_b200.nfield import FieldNeuralCF
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    lr, B, steps = 0.2, 256, 3
    base, _ = _ncf_pair()
    sd = {k: v.detach().clone() for k, v in base.state_dict().items()}

    def batch(r):
        g = torch.Generator().manual_seed(40 + r)
        return (torch.randint(0, 300, (B,), generator=g).cuda(), torch.randint(0, 500, (B,), generator=g).cuda(),
                (torch.rand(B, 1, generator=g) < 0.4).float().cuda())

    def rank_fn(fab):
        m = FieldNeuralCF(300, 500, 16, [32, 16, 8], seed=1, device="cuda", sharded=True, fabric=fab)
        m.gmf.load_global(sd["gmf.weight"])
        m.mlp.load_global(sd["mlp.weight"])
        m.load_state_dict({k: v for k, v in sd.items() if not k.endswith(("gmf.weight", "mlp.weight"))}, strict=False)
        dense = [p for p in m.parameters() if p.requires_grad]
        tr = Trainer(m, torch.nn.BCELoss(), FusedRowOptimizer(m, torch.optim.SGD(dense, lr=lr), lr=lr))
        u, i, y = batch(fab.rank)
        preds = []
        for _ in range(steps):
            tr.train_loop(u, i, train_rating=y)
            preds.append(tr.predictions_train.detach().clone())
        return m.gmf.weight.data.clone(), m.mlp.weight.data.clone(), torch.stack(preds), m.linear2.weight.detach().clone()

    outs = rsdist.ThreadFabric.run(world, rank_fn)
    ops.check_status()
    dense = [p for p in base.parameters() if p.requires_grad]
    tr = Trainer(base, torch.nn.BCELoss(), FusedRowOptimizer(base, torch.optim.SGD(dense, lr=lr), lr=lr))
    bs = [batch(r) for r in range(world)]
    u, i, y = (torch.cat([b[k] for b in bs]) for k in range(3))
    for s in range(steps):
        tr.train_loop(u, i, train_rating=y)
        want = tr.predictions_train.detach().cpu().numpy()
        for r in range(world):
            np.testing.assert_allclose(outs[r][2][s].cpu().numpy(), want[r * B:(r + 1) * B], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rsdist.unshard_rows([o[0] for o in outs]).cpu().numpy(), base.gmf.weight.cpu().numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(rsdist.unshard_rows([o[1] for o in outs]).cpu().numpy(), base.mlp.weight.cpu().numpy(), rtol=1e-5, atol=2e-6)
    for r in range(world):
        np.testing.assert_allclose(outs[r][3].cpu().numpy(), base.linear2.weight.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
