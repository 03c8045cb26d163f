"""Drop-in modules on the B200 against (a) the golden vectors recorded from the unmodified reference and (b) the
CPU oracle on fresh seeded inputs.  Bar: 1e-5 relative (fp32) on predictions, loss, every gradient and the
parameters after two Trainer.train_loop steps with the scripts' Adam(lr=1e-3, weight_decay=1e-5)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import feature_matrix
from oracle import ml100k

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def build(name):
    from deeplearningrecommendationsystem_b200 import model as M
    return {
        "lr": lambda: M.LogisticRegression(50, 60, 43),
        "mf": lambda: M.MatrixFactorization(50, 60, 16),
        "deepfm": lambda: M.DeepFM(50, 60, [32, 16, 8, 1], 16),
        "nfm": lambda: M.NFM(50, 60, [32, 16, 8, 1], 16),
        "afm": lambda: M.AFM(50, 60, 16, 8),
        "ffm": lambda: M.FFM(43, 8),
        "pnn_in": lambda: M.PNN(8, [32, 16, 8, 4], "in"),
        "pnn_out": lambda: M.PNN(16, [32, 16, 8, 4], "out"),
        "din": lambda: M.DIN(60, 16),
        "dien": lambda: M.DIEN(60, 16),
        "neuralcf": lambda: M.NeuralCF(50, 60, 8, [32, 16, 8]),
        "widedeep": lambda: M.WideDeep(50, 60, [32, 16, 8, 1], 16),
        "deepcross": lambda: M.DeepCross(50, 60, 3, [32, 16], 8),
        "deepcrossing": lambda: M.DeepCrossing(50, 60, 8, [32, 16]),
    }[name]()


NAMES = ["lr", "mf", "deepfm", "nfm", "afm", "ffm", "pnn_in", "pnn_out", "din", "dien", "neuralcf",
         "widedeep", "deepcross", "deepcrossing"]


def close(got, want, atol, msg=""):
    np.testing.assert_allclose(got.detach().cpu().numpy(), np.asarray(want), rtol=RTOL, atol=atol, err_msg=msg)


@pytest.mark.parametrize("name", NAMES)
def test_state_dict_is_reference_compatible(name):
    _, _, sd0, _ = load_golden(name)
    m = build(name)
    own = m.state_dict()
    assert list(own.keys()) == list(sd0.keys())
    for k in sd0:
        assert tuple(own[k].shape) == tuple(sd0[k].shape), k
    m.load_state_dict(sd0, strict=True)


@pytest.mark.parametrize("name", NAMES)
def test_forward_loss_grads_match_reference(name):
    ins, y, sd0, z = load_golden(name)
    m = build(name)
    m.load_state_dict(sd0)
    m = m.cuda()
    pred = m(*[t.cuda() for t in ins])
    assert tuple(pred.shape) == tuple(z["pred"].shape)
    loss = torch.nn.BCELoss()(pred, y.cuda())
    loss.backward()
    close(pred, z["pred"], 1e-6)
    np.testing.assert_allclose(loss.item(), z["loss"], rtol=RTOL)
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        close(p.grad, z[f"grad/{k}"], 1e-7, k)
    from deeplearningrecommendationsystem_b200 import ops
    ops.check_status()


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("optim_kind", ["torch", "fused_dense"])
def test_two_trainer_steps_match_reference(name, optim_kind):
    from deeplearningrecommendationsystem_b200.optim import DenseAdam
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    ins, y, sd0, z = load_golden(name)
    m = build(name)
    m.load_state_dict(sd0)
    m = m.cuda()
    opt = (torch.optim.Adam if optim_kind == "torch" else DenseAdam)(m.parameters(), lr=1e-3, weight_decay=1e-5)
    tr = Trainer(m, torch.nn.BCELoss(), opt)
    cin, cy = [t.cuda() for t in ins], y.cuda()
    losses = []
    for _ in range(2):
        tr.train_loop(*cin, train_rating=cy)
        losses.append(tr.train_loss.item())
    np.testing.assert_allclose(losses, z["losses"], rtol=RTOL)
    # Adam divides by sqrt(v): an element whose gradient sits at rounding-noise level moves by up to lr per step
    # in a direction that noise decides, on ANY two fp32 implementations (CPU vs CUDA torch differ the same way).
    # So: at least 99% of every tensor within 1e-5 relative (+2e-6), and no element further than steps*lr.
    for k, v in m.state_dict().items():
        got, want = v.detach().cpu().numpy().astype(np.float64), np.asarray(z[f"sd2/{k}"], dtype=np.float64)
        err = np.abs(got - want)
        ok = err <= 2e-6 + RTOL * np.abs(want)
        assert ok.mean() >= 0.99, f"{k}: only {ok.mean():.4f} of elements within 1e-5"
        assert err.max() <= 2 * 1e-3 * 1.01, f"{k}: max deviation {err.max():.3e} exceeds steps*lr"
    tr.valid_loop(*cin, valid_rating=cy)
    np.testing.assert_allclose(tr.predictions_valid.cpu().numpy(), z["pred_after"], rtol=1e-4, atol=1e-6)


def test_pnn_out_raises_when_batch_differs_from_dim():
    ins, _, sd0, _ = load_golden("pnn_out")
    m = build("pnn_out")
    m.load_state_dict(sd0)
    with pytest.raises(RuntimeError):
        m.cuda()(ins[0][:7].cuda())


@pytest.mark.parametrize("name", ["lr", "deepfm", "nfm", "afm", "ffm", "pnn_in", "widedeep", "deepcross", "deepcrossing"])
def test_against_oracle_on_fresh_inputs(name):
    """larger batch, more duplicates, different seed than the golden fixtures; also run twice for determinism."""
    _, _, sd0, _ = load_golden(name)
    g = torch.Generator().manual_seed(99)
    nu, ni = (943, 1682) if name in ("ffm", "pnn_in") else (50, 60)
    B = 1500
    x = feature_matrix(g, B, nu, ni)
    x[:, 0] = (torch.rand(B, generator=g) ** 3 * nu).floor()        # skewed users -> long duplicate segments
    y = (torch.rand(B, 1, generator=g) < 0.4).float()
    pred, loss, grads = ml100k.loss_and_grads(name, sd0, [x], y)
    runs = []
    for _ in range(2):
        m = build(name)
        m.load_state_dict(sd0)
        m = m.cuda()
        p = m(x.cuda())
        l = torch.nn.BCELoss()(p, y.cuda())
        l.backward()
        runs.append({k: v.grad.clone() for k, v in m.named_parameters()})
        close(p, pred.numpy(), 1e-6)
        np.testing.assert_allclose(l.item(), loss.item(), rtol=RTOL)
        for k, v in m.named_parameters():
            close(v.grad, grads[k].numpy(), 2e-7, k)
    for k in runs[0]:
        if "embed" in k or k in ("user.weight", "item.weight") or k.endswith(("_user.weight", "_item.weight")):
            assert torch.equal(runs[0][k], runs[1][k]), f"{k}: embedding gradient not bitwise reproducible"


@pytest.mark.parametrize("name", ["mf", "neuralcf", "din", "dien"])
def test_id_models_against_oracle_on_fresh_inputs(name):
    _, _, sd0, _ = load_golden(name)
    g = torch.Generator().manual_seed(17)
    B = 1200
    if name in ("mf", "neuralcf"):
        ins = [(torch.rand(B, generator=g) ** 2 * 50).long(), torch.randint(0, 60, (B,), generator=g)]
    else:
        ins = [torch.randint(0, 60, (B, 23), generator=g), torch.randint(0, 60, (B,), generator=g)]
    y = (torch.rand(B, generator=g) < 0.4).float()
    y = y if name == "mf" else y.unsqueeze(1)
    pred, loss, grads = ml100k.loss_and_grads(name, sd0, ins, y)
    m = build(name)
    m.load_state_dict(sd0)
    m = m.cuda()
    p = m(*[t.cuda() for t in ins])
    l = torch.nn.BCELoss()(p, y.cuda())
    l.backward()
    close(p, pred.numpy(), 1e-6)
    np.testing.assert_allclose(l.item(), loss.item(), rtol=RTOL)
    for k, v in m.named_parameters():
        close(v.grad, grads[k].numpy(), 2e-7, k)


def test_eval_mode_and_no_grad():
    ins, y, sd0, z = load_golden("deepfm")
    m = build("deepfm")
    m.load_state_dict(sd0)
    m = m.cuda().eval()
    with torch.no_grad():
        p = m(ins[0].cuda())
    assert not p.requires_grad
    close(p, z["pred"], 1e-6)
    assert next(m.parameters()).device.type == "cuda"


@pytest.mark.parametrize("pool", [True, False])
def test_din_attention_tc_equals_fused(pool):
    """the tcgen05 kernels (what din_attention picks for B*L >= 8192 rows) and the CUDA-core kernel pair (RS_DIN_TC=0)
    agree on outputs and gradients through autograd."""
    import os
    from deeplearningrecommendationsystem_b200 import attention, ops
    g = torch.Generator().manual_seed(0)
    B, L, D = 90, 100, 64
    unit = torch.nn.Sequential(torch.nn.Linear(3 * D, 128), torch.nn.ReLU(), torch.nn.Linear(128, 64), torch.nn.ReLU(),
                               torch.nn.Linear(64, 1)).cuda()
    rows = (torch.randn(B, L + 1, D, generator=g) * 0.5).cuda()
    gup = torch.randn((B, D) if pool else (B, L, D), generator=g).cuda()
    res = {}
    for impl in ("tc", "fused"):
        if impl == "fused":
            os.environ["RS_DIN_TC"] = "0"
        try:
            r = rows.clone().requires_grad_(True)
            unit.zero_grad()
            n0 = ops.launches()
            out = attention.din_attention(r, unit, pool)
            (out * gup).sum().backward()
            res[impl] = [out.detach(), r.grad] + [p.grad.clone() for p in unit.parameters()]
            res[impl + "_launches"] = ops.launches() - n0
        finally:
            os.environ.pop("RS_DIN_TC", None)
    assert res["tc_launches"] > res["fused_launches"]                      # two different kernel sets really ran
    for k, (a, b) in enumerate(list(zip(res["tc"], res["fused"]))[:-1]):   # the last entry is d/d b2, analytically zero
        tol = 1e-5 * max(1e-3, float(b.abs().max()))
        if k == 1:   # d rows: a ReLU input within rounding of zero may flip between the two evaluations (see kernel tests)
            bad = int(((a - b).abs() > tol + 1e-5 * b.abs()).flatten(1).any(dim=1).sum())
            assert bad <= 1, f"{bad} samples differ"
        else:
            assert float((a - b).norm()) <= 1e-4 * float(b.norm()) + tol, k
    assert float(res["tc"][-1].abs().max()) < 1e-4 and float(res["fused"][-1].abs().max()) < 1e-4


def test_c1_shape_deepfm_matches_oracle():
    """BASELINE.json configs[0] at its own shape (scripts/deepfm.py:52-55): DeepFM(943, 1682, [512, 256, 128, 1], 128) on one
    full batch of 87 909 ML-100k-shaped rows, dense Adam(1e-3, weight_decay=1e-5) through Trainer.train_loop.  Predictions,
    loss and every gradient against the CPU oracle (oracle/ml100k.py, pinned to the reference's golden outputs at the
    small shape); then the second step's loss, which only agrees if the first Adam step did.
    Bar: 1e-5 relative; gradients relative to the largest entry of their tensor (a full-batch sum of 87 909 terms)."""
    from helpers import feature_matrix_fast
    from deeplearningrecommendationsystem_b200 import model as M
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    from oracle import optim as oo
    g = torch.Generator().manual_seed(7)
    B = 87909
    x = feature_matrix_fast(g, B, 943, 1682)
    y = (torch.rand(B, 1, generator=g) < 0.3).float()
    torch.manual_seed(0)
    m = M.DeepFM(943, 1682, [512, 256, 128, 1], 128)
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    pred, loss, grads = ml100k.loss_and_grads("deepfm", sd0, [x], y)
    m = m.cuda()
    tr = Trainer(m, torch.nn.BCELoss(), torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5))
    tr.train_loop(x.cuda(), train_rating=y.cuda())
    close(tr.predictions_train, pred.numpy(), 1e-6)
    np.testing.assert_allclose(tr.train_loss.item(), loss.item(), rtol=RTOL)
    for k, v in m.named_parameters():
        want = grads[k].numpy()
        np.testing.assert_allclose(v.grad.cpu().numpy(), want, rtol=RTOL, atol=RTOL * float(np.abs(want).max()) + 1e-9, err_msg=k)
    # oracle Adam step (torch.optim.Adam single-tensor arithmetic, oracle/optim.adam_dense), then the second forward
    sd1 = {}
    for k, p in sd0.items():
        sd1[k] = oo.adam_dense(p, grads[k], torch.zeros_like(p), torch.zeros_like(p), 1, lr=1e-3, wd=1e-5)[0]
    _, loss1, _ = ml100k.loss_and_grads("deepfm", sd1, [x], y)
    tr.train_loop(x.cuda(), train_rating=y.cuda())
    np.testing.assert_allclose(tr.train_loss.item(), loss1.item(), rtol=RTOL)
