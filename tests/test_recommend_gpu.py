"""Catalogue ranking on the B200: rs_rank_segments / rs_mf_rank against the numpy oracle, and the
drop-in modules' recommendation() against what the unmodified reference returned (tests/golden/recommend.npz).

Index work: bit-exact against the oracle for the same fp32 scores.  Against the reference, scores come from a
different fp32 GEMM, so two items may swap only where the reference's own scores are closer than 1e-5 relative."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import GOLDEN, load_golden
from oracle import ranking as R

pytestmark = pytest.mark.gpu


def _ops():
    from deeplearningrecommendationsystem_b200 import ops
    return ops


def _scores(n, seed, ties=False):
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(n, generator=g)
    if ties:
        s = (s * 4).round() / 4                      # heavy ties
        s[::11] = 0.0
        s[5::13] = -0.0
        s[7::29] = float("nan")
        s[3::31] = float("inf")
        s[2::37] = float("-inf")
    return s


@pytest.mark.parametrize("seg_len,k", [(1, 1), (7, 3), (64, 64), (100, 100), (1682, 1682), (1682, 50), (2048, 2048),
                                       (5000, 17), (16384, 16384)])
@pytest.mark.parametrize("ties", [False, True])
def test_uniform_segments_match_oracle(seg_len, k, ties):
    S = 5 if seg_len > 4096 else 37
    s = _scores(S * seg_len, seg_len + k, ties)
    idx, val = _ops().rank_segments(s.cuda(), k, seg_len=seg_len, values=True)
    want = R.rank_segments(s.numpy(), np.arange(S + 1) * seg_len, k)
    assert np.array_equal(idx.cpu().numpy(), want)
    got_val = val.cpu().numpy()
    ref_val = np.take_along_axis(s.numpy().reshape(S, seg_len), want, axis=1)
    assert np.array_equal(np.isnan(got_val), np.isnan(ref_val))
    assert np.array_equal(np.nan_to_num(got_val, nan=0.0) + 0.0, np.nan_to_num(ref_val, nan=0.0) + 0.0)


def test_ragged_segments_match_oracle():
    g = torch.Generator().manual_seed(3)
    lens = torch.randint(9, 700, (61,), generator=g)
    lens[7] = 9
    seg = torch.zeros(62, dtype=torch.int64)
    seg[1:] = torch.cumsum(lens, 0)
    s = _scores(int(seg[-1]), 5, ties=True)
    idx = _ops().rank_segments(s.cuda(), 9, seg_start=seg.cuda())
    assert np.array_equal(idx.cpu().numpy(), R.rank_segments(s.numpy(), seg.numpy(), 9))


@pytest.mark.parametrize("seg_len,k", [(40000, 100), (100001, 8192), (16385, 1)])
def test_long_segments_are_streamed(seg_len, k):
    s = _scores(3 * seg_len, seg_len, ties=(k == 100))
    idx = _ops().rank_segments(s.cuda(), k, seg_len=seg_len)
    assert np.array_equal(idx.cpu().numpy(), R.rank_segments(s.numpy(), np.arange(4) * seg_len, k))


def test_k_out_of_range_raises_like_topk():
    ops = _ops()
    s = _scores(100, 1).cuda()
    with pytest.raises(RuntimeError, match="out of range"):
        ops.rank_segments(s, 11, seg_len=10)
    seg = torch.tensor([0, 50, 55, 100]).cuda()
    with pytest.raises(RuntimeError, match="out of range"):
        ops.rank_segments(s, 6, seg_start=seg)
    with pytest.raises(RuntimeError, match="longer than"):
        ops.rank_segments(s, 64, seg_start=torch.tensor([0, 100]).cuda(), max_len=64)
    with pytest.raises(RuntimeError, match="rs_rank_segments"):            # 20000 > 16384 slots and k > 8192
        ops.rank_segments(_scores(40000, 2).cuda(), 9000, seg_len=20000)
    assert ops.rank_segments(torch.empty(0, device="cuda"), 3, seg_start=torch.zeros(1, dtype=torch.int64).cuda()).shape == (0, 3)


@pytest.mark.parametrize("nu,ni,W,k", [(50, 60, 16, 60), (943, 1682, 64, 1682), (33, 1682, 64, 50), (4, 20000, 32, 100), (9, 130, 10, 7),
                                       # the tiled kernel (>= 64 users, k <= 512 << items): ragged user tile, padded item chunk
                                       (300, 5000, 32, 10), (65, 4100, 64, 100), (203, 20001, 64, 512), (64, 70000, 16, 1)])
def test_mf_rank_matches_oracle(nu, ni, W, k):
    g = torch.Generator().manual_seed(nu + ni)
    U, V = torch.randn(nu, W, generator=g) * 0.3, torch.randn(ni, W, generator=g) * 0.3
    idx, val = _ops().mf_rank(U.cuda(), V.cuda(), k, values=True)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    exact = R.mf_scores(U.numpy(), V.numpy())
    # the returned scores are the fp32 dot products ...
    np.testing.assert_allclose(val, np.take_along_axis(exact, idx, axis=1), rtol=1e-5, atol=1e-6)
    # ... and the ranking is exactly the oracle's ranking of those scores (every row is a permutation prefix)
    assert (np.diff(val, axis=1) <= 0).all()
    for u in range(nu):
        assert len(set(idx[u].tolist())) == k
        kth = val[u, -1]
        assert (np.delete(exact[u], idx[u]) <= kth + 1e-5).all()             # nothing better was left out


def _same_ranking(got, want, scores, what):
    """identical, except that items whose reference scores are within 1e-5 relative may swap."""
    assert got.shape == want.shape, what
    for u in range(want.shape[0]):
        bad = np.flatnonzero(got[u] != want[u])
        for r in bad:
            a, b = scores[u][got[u][r]], scores[u][want[u][r]]
            assert abs(a - b) <= 1e-5 * max(abs(a), abs(b)) + 1e-7, f"{what}: user {u} rank {r}: {got[u][r]} vs {want[u][r]}"


def _frame(values):
    cols = ["user_id", "item_id", "age"] + [f"c{k}" for k in range(values.shape[1] - 3)]
    df = pd.DataFrame(values, columns=cols)
    df["user_id"] = df["user_id"].astype(np.int64)
    df["item_id"] = df["item_id"].astype(np.int64)
    return df


def _model(name):
    from test_models_gpu import build
    _, _, sd0, _ = load_golden(name)
    m = build(name)
    m.load_state_dict(sd0)
    return m.cuda().eval()


def test_recommendation_matches_reference():
    z = np.load(os.path.join(GOLDEN, "recommend.npz"))
    _same_ranking(_model("mf").recommendation(50, 60), z["mf/idx"], z["mf/scores"], "mf")
    df = _frame(z["frame"])
    for name, k in (("deepfm", 20), ("widedeep", 7), ("pnn_in", 20)):
        _same_ranking(_model(name).recommendation(12, df, k), z[f"{name}/idx"], z[f"{name}/scores"], name)
    _same_ranking(_model("pnn_out").recommendation(5, _frame(z["frame_out"]), 16), z["pnn_out/idx"], z["pnn_out/scores"], "pnn_out")
    _same_ranking(_model("neuralcf").recommendation(50, 60), z["neuralcf/idx"], z["neuralcf/scores"], "neuralcf")
    hist = z["hist"].tolist()
    for name in ("din", "dien"):
        _same_ranking(_model(name).recommendation(9, 60, hist, 10), z[f"{name}/idx"], z[f"{name}/scores"], name)


def test_recommendation_full_catalogue_scale():
    """943 x 1682 (the scripts' call, scripts/deepfm.py:67 with k = num_items): chunked forward + one ranking launch
    must equal ranking each user's scores separately."""
    from helpers import catalogue_frame
    from deeplearningrecommendationsystem_b200 import model as M
    torch.manual_seed(0)
    m = M.DeepFM(943, 1682, [64, 32, 1], 16).cuda().eval()
    g = torch.Generator().manual_seed(8)
    nu, ni = 160, 1682                                     # 269k rows: two forward chunks
    df = catalogue_frame(g, nu, ni)
    got = m.recommendation(nu, df, ni)
    assert got.shape == (nu, ni) and got.dtype == np.int64
    with torch.no_grad():
        for u in (0, 57, 159):
            rows = torch.tensor(df[df["user_id"] == u].values, dtype=torch.float32).cuda()
            s = m(rows).reshape(-1).cpu().numpy()
            _same_ranking(got[u:u + 1], R.rank_desc(s, ni)[None], [s], f"user {u}")
    mf = M.MatrixFactorization(943, 1682, 64).cuda()
    r = mf.recommendation(943, 1682)
    assert r.shape == (943, 1682) and (np.sort(r, axis=1) == np.arange(1682)).all()
