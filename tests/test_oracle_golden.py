"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/*.npz)."""
import random

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ml100k, optim, index, ranking, sampling, interactions as I

MODELS = ["lr", "mf", "deepfm", "nfm", "afm", "ffm", "pnn_in", "pnn_out", "din", "dien", "neuralcf",
          "widedeep", "deepcross", "deepcrossing"]
RTOL, ATOL = 1e-5, 1e-6       # north star: 1e-5 relative (fp32) for logits, loss, gradients


@pytest.mark.parametrize("name", MODELS)
def test_forward_loss_grads(name):
    ins, y, sd0, z = load_golden(name)
    pred, loss, grads = ml100k.loss_and_grads(name, sd0, ins, y)
    assert pred.shape == z["pred"].shape            # MF is (B,), everything else (B,1)
    np.testing.assert_allclose(pred.numpy(), z["pred"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(loss.item(), z["loss"], rtol=RTOL)
    for k in sd0:
        np.testing.assert_allclose(grads[k].numpy(), z[f"grad/{k}"], rtol=RTOL, atol=ATOL, err_msg=k)


@pytest.mark.parametrize("name", MODELS)
def test_two_adam_steps(name):
    """oracle fwd/bwd + oracle dense Adam == reference Trainer.train_loop x2 with Adam(1e-3, wd=1e-5)."""
    ins, y, sd0, z = load_golden(name)
    sd = {k: v.clone() for k, v in sd0.items()}
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v = {k: torch.zeros_like(t) for k, t in sd.items()}
    losses = []
    for step in (1, 2):
        _, loss, g = ml100k.loss_and_grads(name, sd, ins, y)
        losses.append(loss.item())
        for k in sd:
            sd[k], m[k], v[k] = optim.adam_dense(sd[k], g[k], m[k], v[k], step, lr=1e-3, wd=1e-5)
    np.testing.assert_allclose(losses, z["losses"], rtol=RTOL)
    for k in sd:
        np.testing.assert_allclose(sd[k].numpy(), z[f"sd2/{k}"], rtol=1e-5, atol=2e-6, err_msg=k)
    with torch.no_grad():
        np.testing.assert_allclose(ml100k.forward(name, sd, *ins).numpy(), z["pred_after"], rtol=1e-4, atol=1e-6)


def test_pnn_out_requires_b_eq_d():
    ins, y, sd0, _ = load_golden("pnn_out")
    with pytest.raises(RuntimeError):
        ml100k.forward("pnn_out", sd0, ins[0][:7])


def test_sampler_stream():
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "sampler.npz"))
    excl = {tuple(p) for p in z["excl"].tolist()}
    random.seed(123)
    u1, i1 = index.negative_draws(20, 30, excl, 5)
    u2, i2 = index.negative_draws(20, 30, excl, 2)
    assert u1 == z["u1"].tolist() and i1 == z["i1"].tolist()
    assert u1 + u2 == z["u2"].tolist() and i1 + i2 == z["i2"].tolist()      # state accumulates
    random.seed(321)
    u3, i3 = index.negative_draws(20, 30, excl, 3)
    assert u3 == z["df_user"].tolist() and i3 == z["df_item"].tolist()


def test_dedup_matches_torch_unique():
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 50, (1000,), generator=g)
    u, inv, cnt = index.dedup(ids.numpy())
    tu, tinv, tcnt = torch.unique(ids, sorted=True, return_inverse=True, return_counts=True)
    assert np.array_equal(u, tu.numpy()) and np.array_equal(inv, tinv.numpy()) and np.array_equal(cnt, tcnt.numpy())
    order = index.stable_order(ids.numpy())
    assert np.array_equal(ids.numpy()[order], np.sort(ids.numpy()))


def test_interaction_identities():
    """FM sum-square == sum of pairwise dots; bi-interaction summed over d == FM; fast FFM == loop FFM."""
    torch.manual_seed(0)
    E = torch.randn(5, 7, 8)
    np.testing.assert_allclose(I.fm_second_order(E).numpy(), I.inner_products(E).sum(1).numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(I.bi_interaction(E).sum(1).numpy(), I.fm_second_order(E).numpy(), rtol=1e-4, atol=1e-5)
    T = torch.randn(4, 6, 6, 4)
    np.testing.assert_allclose(I.ffm_cross(T).numpy(), I.ffm_cross_fast(T).numpy(), rtol=1e-5, atol=1e-6)


def test_sparse_sgd_equals_dense_sgd():
    torch.manual_seed(1)
    table = torch.randn(20, 4)
    ids = torch.randint(0, 20, (64,))
    G = torch.randn(64, 4)
    dense = torch.zeros_like(table).index_add_(0, ids, G)
    want = table - 0.1 * dense
    got = optim.sgd_rows(table.clone(), ids, G, 0.1)
    assert torch.equal(got, want)


# ---- catalogue ranking and the evaluator metrics
def _npz(name):
    import os
    from conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, name))


def _unragged(flat, off):
    return [flat[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


@pytest.mark.parametrize("name", ["mf", "deepfm", "widedeep", "pnn_in", "pnn_out", "neuralcf", "din", "dien"])
def test_rank_oracle_equals_reference_topk(name):
    """rank_desc(scores) reproduces what the reference's recommendation() returned for the same scores."""
    z = _npz("recommend.npz")
    idx, scores = z[f"{name}/idx"], z[f"{name}/scores"]
    for u in range(idx.shape[0]):
        assert ranking.rank_desc(scores[u], idx.shape[1]).tolist() == idx[u].tolist(), (name, u)


def test_mf_scores_oracle():
    _, _, sd0, _ = load_golden("mf")
    z = _npz("recommend.npz")
    got = ranking.mf_scores(sd0["user_embeddings.weight"].numpy(), sd0["item_embeddings.weight"].numpy())
    np.testing.assert_allclose(got, z["mf/scores"], rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("k", [5, 50, 200])
def test_ranking_metrics_oracle_and_mirror(k):
    from deeplearningrecommendationsystem_b200.evaluator import Ranking
    z = _npz("ranking.npz")
    actual, predicted = _unragged(z["actual"], z["actual_off"]), _unragged(z["predicted"], z["predicted_off"])
    np.testing.assert_allclose(ranking.ranking_metrics(actual, predicted, k), z[f"k{k}"], rtol=1e-12)
    r = Ranking(actual, predicted, k)
    got = [*r.precision_recall_f1(), r.mapk(), r.mean_ndcg(), r.mrr()]
    np.testing.assert_allclose(got, z[f"k{k}"], rtol=1e-14)
    assert Ranking.apk(actual[5], predicted[5], k) == 0.0 and Ranking.rr(actual[5], predicted[5]) == 0.0   # no hit at all
    r.ranking_eval()


def test_binary_metrics_oracle_and_mirror():
    from deeplearningrecommendationsystem_b200.evaluator import Evaluator
    z = _npz("evaluator.npz")
    np.testing.assert_allclose(ranking.binary_metrics(z["y_true"], z["y_pred"]), z["metrics"], rtol=1e-12)
    got = Evaluator.eval(torch.from_numpy(z["y_true"]), torch.from_numpy(z["y_pred"]))
    np.testing.assert_allclose(got, z["metrics"], rtol=1e-12)


# ---- device-side input preparation
def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    def run(ctr, key):
        return [int(v) for v in sampling.philox4x32_10(np.array(ctr, np.uint32), np.array(key, np.uint32))]
    assert run([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_negative_draws_follow_the_reference_rule():
    z = _npz("sampler.npz")
    excl = {tuple(p) for p in z["excl"].tolist()}
    u, i = sampling.negative_draws_philox(20, 30, excl, 5, seed=9)
    assert u.tolist() == np.repeat(np.arange(20), 5).tolist()                    # user-major, 5 per user (sampler.py:21-22)
    assert all((a, b) not in excl for a, b in zip(u.tolist(), i.tolist()))       # never an observed pair (sampler.py:24-25)
    assert i.min() >= 0 and i.max() < 30
    u2, i2 = sampling.negative_draws_philox(20, 30, excl, 5, seed=9)
    assert i.tolist() == i2.tolist()
    assert i.tolist() != sampling.negative_draws_philox(20, 30, excl, 5, seed=9, epoch=1)[1].tolist()
    # uniform over the free items: chi-square against the exact per-user free sets, 400 draws per user
    u, i = sampling.negative_draws_philox(20, 30, excl, 400, seed=3)
    chi = 0.0
    dof = 0
    for user in range(20):
        free = [it for it in range(30) if (user, it) not in excl]
        cnt = np.bincount(i[u == user], minlength=30)[free]
        e = 400 / len(free)
        chi += float(((cnt - e) ** 2 / e).sum())
        dof += len(free) - 1
    assert abs(chi - dof) < 5 * np.sqrt(2 * dof)


def test_assemble_features_oracle_equals_reference_merge():
    z = _npz("features.npz")
    got = sampling.assemble_features(z["users"], z["items"], z["user_feat"], z["item_feat"])
    assert np.array_equal(got, z["x"])                    # bit-exact, row order of the input pairs
    assert np.array_equal(z["rating"], z["pair_rating"])


def test_rank_oracle_tie_and_nan_conventions():
    """distinct scores: exactly torch.topk; ties: the lower position first (torch.sort(stable=True, descending=True));
    NaN ranks above everything, as in torch.topk."""
    g = torch.Generator().manual_seed(1)
    s = torch.randn(500, generator=g)
    assert ranking.rank_desc(s.numpy(), 37).tolist() == torch.topk(s, 37).indices.tolist()
    t = (s * 2).round() / 2                                    # heavy ties
    assert ranking.rank_desc(t.numpy(), 500).tolist() == torch.sort(t, stable=True, descending=True).indices.tolist()
    t[17] = float("nan")
    assert int(ranking.rank_desc(t.numpy(), 1)[0]) == 17 == int(torch.topk(t, 1).indices[0])
    with pytest.raises(RuntimeError):
        ranking.rank_desc(s.numpy()[:5], 6)
