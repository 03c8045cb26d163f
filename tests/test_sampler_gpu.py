"""Device-side input preparation on the B200: the optional Philox sampler against its numpy
restatement (bit-exact) and the feature-matrix assembly against the reference's own pd.merge output (bit-exact)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import sampling

pytestmark = pytest.mark.gpu


def _excl(nu, ni, n, seed):
    rng = np.random.default_rng(seed)
    return {(int(a), int(b)) for a, b in zip(rng.integers(0, nu, n), rng.integers(0, ni, n))}


@pytest.mark.parametrize("nu,ni,n_excl,per_user,seed", [(20, 30, 150, 5, 9), (943, 1682, 100000, 30, 2 ** 40 + 17), (7, 3, 10, 4, 1),
                                                        (5, 1000, 0, 3, 0)])
def test_device_stream_equals_oracle(nu, ni, n_excl, per_user, seed):
    from deeplearningrecommendationsystem_b200.sampler import DeviceSampler
    excl = _excl(nu, ni, n_excl, seed % 1000)
    if nu == 7:                                      # leave every user at least one free item
        excl = {p for p in excl if p[1] != p[0] % 3}
    s = DeviceSampler(seed=seed)
    u, i, r = s.negative_sampling(nu, ni, excl, per_user)
    wu, wi = sampling.negative_draws_philox(nu, ni, excl, per_user, seed)
    assert u.dtype == torch.int64 and r.dtype == torch.float32 and float(r.abs().sum()) == 0.0
    assert np.array_equal(u.cpu().numpy(), wu) and np.array_equal(i.cpu().numpy(), wi)
    assert all((a, b) not in excl for a, b in zip(wu.tolist(), i.cpu().tolist()))
    # second call: a fresh epoch, appended after the first (sampler/sampler.py:13-14,26-27 accumulate on the instance)
    u2, i2, r2 = s.negative_sampling(nu, ni, excl, per_user)
    _, wi2 = sampling.negative_draws_philox(nu, ni, excl, per_user, seed, epoch=1)
    assert u2.numel() == 2 * nu * per_user == r2.numel()
    assert np.array_equal(i2.cpu().numpy(), np.concatenate([wi, wi2]))


def test_saturated_user_is_reported():
    from deeplearningrecommendationsystem_b200.sampler import DeviceSampler
    excl = {(0, it) for it in range(50)}                                  # user 0 has seen everything
    with pytest.raises(RuntimeError, match="no free item"):
        DeviceSampler(seed=1).negative_sampling(2, 50, excl, 1)


def test_assemble_features_matches_reference_merge():
    from deeplearningrecommendationsystem_b200 import ops
    from deeplearningrecommendationsystem_b200.sampler import DeviceSampler
    z = np.load(os.path.join(GOLDEN, "features.npz"))
    t = lambda k: torch.from_numpy(z[k]).cuda()
    x = DeviceSampler.features(t("users"), t("items"), t("user_feat"), t("item_feat"))
    assert x.dtype == torch.float32 and np.array_equal(x.cpu().numpy(), z["x"])
    ops.check_status()
    assert DeviceSampler.features(t("users")[:0], t("items")[:0], t("user_feat"), t("item_feat")).shape == (0, 45)
    bad = t("users").clone()
    bad[3] = 10 ** 6
    DeviceSampler.features(bad, t("items"), t("user_feat"), t("item_feat"))
    with pytest.raises(IndexError):
        ops.check_status()


def test_sampled_batch_trains_a_drop_in_model():
    """sampler -> features -> DeepFM step, all on the device: the path scripts/deepfm.py:20-49 prepares on the host."""
    from deeplearningrecommendationsystem_b200 import model as M
    from deeplearningrecommendationsystem_b200.sampler import DeviceSampler
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    z = np.load(os.path.join(GOLDEN, "features.npz"))
    nu, ni = z["user_feat"].shape[0], z["item_feat"].shape[0]
    pos = {(int(a), int(b)) for a, b in zip(z["users"], z["items"])}
    s = DeviceSampler(seed=5)
    nusr, nitm, nr = s.negative_sampling(nu, ni, pos, 4)
    pu, pi = torch.from_numpy(z["users"]).cuda(), torch.from_numpy(z["items"]).cuda()
    users, items = torch.cat([pu, nusr]), torch.cat([pi, nitm])
    rating = torch.cat([torch.ones(pu.numel(), device="cuda"), nr]).unsqueeze(1)
    x = s.features(users, items, torch.from_numpy(z["user_feat"]).cuda(), torch.from_numpy(z["item_feat"]).cuda())
    torch.manual_seed(0)
    m = M.DeepFM(nu, ni, [32, 16, 1], 8).cuda()
    tr = Trainer(m, torch.nn.BCELoss(), torch.optim.Adam(m.parameters(), lr=1e-2))
    losses = []
    for _ in range(30):
        tr.train_loop(x, train_rating=rating)
        losses.append(tr.train_loss.item())
    assert losses[-1] < losses[0] * 0.9
