"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares;
the host mirrors (Trainer, Sampler) behave like the reference's without a GPU."""
import ctypes
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from deeplearningrecommendationsystem_b200 import _lib


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/recsys_b200.h but not exported"
    assert set(names) - {"rs_last_error"} == set(_lib.SIGNATURES), "ctypes signatures out of sync with the header"
    assert lib.rs_version() >= 100


def test_argument_errors_are_reported_not_thrown():
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    assert lib.rs_dedup_workspace_bytes(-5, 1, ctypes.byref(nbytes)) == -1        # RS_E_ARG
    assert b"rs_dedup_workspace_bytes" in lib.rs_last_error()
    assert lib.rs_dedup_workspace_bytes(1000, 16, ctypes.byref(nbytes)) == 0 and nbytes.value > 1000 * 4 * 10


def test_ops_refuse_cpu_tensors():
    from deeplearningrecommendationsystem_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gather_rows(ops.dummy_tables(1, 4), torch.zeros(3, 1, dtype=torch.int64))


def test_sampler_replays_reference_stream():
    from deeplearningrecommendationsystem_b200.sampler import Sampler
    z = np.load(f"{GOLDEN}/sampler.npz")
    excl = {tuple(p) for p in z["excl"].tolist()}
    random.seed(123)
    s = Sampler()
    u, i, r = s.negative_sampling(20, 30, excl, 5)
    assert u.dtype == torch.int64 and r.dtype == torch.float32
    assert torch.equal(u, torch.from_numpy(z["u1"])) and torch.equal(i, torch.from_numpy(z["i1"])) and torch.equal(r, torch.from_numpy(z["r1"]))
    u, i, r = s.negative_sampling(20, 30, excl, 2)          # accumulates on the instance like the reference
    assert torch.equal(u, torch.from_numpy(z["u2"])) and torch.equal(i, torch.from_numpy(z["i2"])) and r.numel() == 140
    random.seed(321)
    df = Sampler().negative_sampling2(20, 30, excl, 3)
    assert list(df.columns) == ["user_id", "item_id", "rating"]
    assert df["user_id"].tolist() == z["df_user"].tolist() and df["item_id"].tolist() == z["df_item"].tolist()
    assert (df["rating"] == 0).all()


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.tensor([0.5]))

    def forward(self, a, b=None):
        z = a.float().sum(dim=1, keepdim=True) if b is None else (a.float() + b.float()).unsqueeze(1)
        return torch.sigmoid(self.w * z)


def test_trainer_protocol():
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    m = _Toy()
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    tr = Trainer(m, torch.nn.BCELoss(), opt)
    x, y = torch.randn(8, 3), torch.randint(0, 2, (8, 1)).float()
    w0 = m.w.item()
    tr.train_loop(x, train_rating=y)
    assert m.w.item() != w0 and tr.predictions_train.shape == (8, 1) and tr.train_rating is y
    tr.valid_loop(x, valid_rating=y)
    tr.test_loop(torch.arange(8), torch.arange(8), test_rating=y)
    assert not tr.predictions_valid.requires_grad and tr.test_loss.dim() == 0
    with pytest.raises(ValueError):
        tr.train_loop(x, x, x, train_rating=y)
    with pytest.raises(ValueError):
        tr.valid_loop(valid_rating=y)
    mets = tr.metrics()
    assert set(mets) == {"train", "valid", "test"} and all(len(v) == 5 for v in mets.values())
    tr.model_eval(0)


def test_trainer_metrics_match_sklearn():
    from sklearn.metrics import accuracy_score, f1_score, precision_score, recall_score, roc_auc_score
    from deeplearningrecommendationsystem_b200.trainer.trainer import _binary_metrics
    g = torch.Generator().manual_seed(0)
    y = (torch.rand(500, 1, generator=g) < 0.4).float()
    p = torch.rand(500, 1, generator=g)
    got = _binary_metrics(y, p)
    yt, yp = y.numpy().ravel(), (p.numpy().ravel() >= 0.5).astype(int)      # evaluator/evaluator.py:17-19
    want = [accuracy_score(yt, yp), precision_score(yt, yp), recall_score(yt, yp), f1_score(yt, yp), roc_auc_score(yt, yp)]
    np.testing.assert_allclose(got, want, rtol=1e-12)
