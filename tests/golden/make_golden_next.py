#!/usr/bin/env python
"""Golden fixtures for the SURVEY.md section 8(f) rows, from the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_next.py

  widedeep / deepcross / deepcrossing .npz   same record layout as make_golden.py's models
  recommend.npz    recommendation() outputs (reference model/mf.py:28-35, model/deepfm.py:85-95, model/pnn.py:133-143,
                   model/neuralcf.py:61-72, model/din.py:55-66) on a small seeded catalogue, with the per-user scores so
                   that tests can tell a real ranking difference from an fp32 near-tie
  ranking.npz      evaluator/ranking.py metrics on seeded ragged lists
  evaluator.npz    evaluator/evaluator.py metrics (sklearn) on seeded labels / probabilities incl. values of exactly 0.5
  features.npz     data/reader.py:98-101 MovieLens100K.feature (the two pd.merge calls) on seeded side-feature tables
"""
import os
import sys

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
import make_golden as MG  # noqa: E402  (sets sys.path for the reference and tests/helpers.py)
from helpers import feature_matrix, catalogue_frame  # noqa: E402

from model.widedeep import WideDeep  # noqa: E402
from model.deepcross import DeepCross  # noqa: E402
from model.deepcrossing import DeepCrossing  # noqa: E402
from evaluator.ranking import Ranking  # noqa: E402
from evaluator.evaluator import Evaluator  # noqa: E402


def ref_model(name, ctor):
    z = np.load(os.path.join(OUT, f"{name}.npz"))
    m = ctor()
    m.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")})
    return m.eval()


def ragged(lists):
    flat = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists]) if lists else np.zeros(0, np.int64)
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum([len(l) for l in lists], out=off[1:])
    return flat, off


def main():
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(4242)
    B, NU, NI = 96, 50, 60
    x = feature_matrix(g, B, NU, NI)
    y = (torch.rand(B, 1, generator=g) < 0.4).float()
    MG.run("widedeep", lambda: WideDeep(NU, NI, [32, 16, 8, 1], 16), [x], y)
    MG.run("deepcross", lambda: DeepCross(NU, NI, 3, [32, 16], 8), [x], y)
    MG.run("deepcrossing", lambda: DeepCrossing(NU, NI, 8, [32, 16]), [x], y)

    # ---- recommendation(): small catalogues, golden initial weights
    rec = {}
    with torch.no_grad():
        mf = ref_model("mf", lambda: MG.MatrixFactorization(50, 60, 16))
        rec["mf/idx"] = mf.recommendation(50, 60)
        rec["mf/scores"] = (mf.user_embeddings.weight @ mf.item_embeddings.weight.T).numpy()

        nu, ni = 12, 20
        df = catalogue_frame(g, nu, ni)
        rec["frame"] = df.values.astype(np.float64)
        for name, ctor, k in (("deepfm", lambda: MG.DeepFM(50, 60, [32, 16, 8, 1], 16), ni),
                              ("widedeep", lambda: WideDeep(50, 60, [32, 16, 8, 1], 16), 7),
                              ("pnn_in", lambda: MG.PNN(8, [32, 16, 8, 4], "in"), ni)):
            m = ref_model(name, ctor)
            rec[f"{name}/idx"] = m.recommendation(nu, df, k)
            rec[f"{name}/scores"] = np.stack([m(torch.Tensor(df[df["user_id"] == u].values)).numpy().reshape(-1) for u in range(nu)])
        # PNN "out" only broadcasts when a user's row count equals embed_dim (16 here)
        dfo = catalogue_frame(g, 5, 16)
        rec["frame_out"] = dfo.values.astype(np.float64)
        m = ref_model("pnn_out", lambda: MG.PNN(16, [32, 16, 8, 4], "out"))
        rec["pnn_out/idx"] = m.recommendation(5, dfo, 16)
        rec["pnn_out/scores"] = np.stack([m(torch.Tensor(dfo[dfo["user_id"] == u].values)).numpy().reshape(-1) for u in range(5)])

        ncf = ref_model("neuralcf", lambda: MG.NeuralCF(50, 60, 8, [32, 16, 8]))
        rec["neuralcf/idx"] = ncf.recommendation(50, 60)
        items = torch.arange(60)
        rec["neuralcf/scores"] = np.stack([ncf(torch.full((60,), u), items).numpy().reshape(-1) for u in range(50)])

        hist = torch.randint(0, 60, (9, 7), generator=g)
        hist[:, :2] = 0
        rec["hist"] = hist.numpy()
        for name, ctor in (("din", lambda: MG.DIN(60, 16)), ("dien", lambda: MG.DIEN(60, 16))):
            m = ref_model(name, ctor)
            rec[f"{name}/idx"] = m.recommendation(9, 60, hist.tolist(), 10)
            rec[f"{name}/scores"] = np.stack([m(hist[u].repeat(60, 1), items).numpy().reshape(-1) for u in range(9)])
    np.savez_compressed(os.path.join(OUT, "recommend.npz"), **rec)
    print("recommend:", {k: v.shape for k, v in rec.items() if k.endswith("idx")})

    # ---- ranking metrics on ragged lists (lists shrink the way data.remove_itemid leaves them)
    rng = np.random.default_rng(11)
    nusers, nitems = 40, 120
    actual = [rng.choice(nitems, size=rng.integers(1, 25), replace=False).tolist() for _ in range(nusers)]
    actual[3] = actual[3] + actual[3][:2]                 # duplicates in `actual` (AP divides by the raw length)
    predicted = []
    for u in range(nusers):
        p = rng.permutation(nitems)
        predicted.append(p[: rng.integers(60, nitems)].tolist())
    predicted[5] = [i for i in predicted[5] if i not in set(actual[5])]      # a user with no hit at all
    out = {}
    out["actual"], out["actual_off"] = ragged(actual)
    out["predicted"], out["predicted_off"] = ragged(predicted)
    for k in (5, 50, 200):
        r = Ranking(actual, predicted, k)
        pr, rc, f1 = r.precision_recall_f1()
        out[f"k{k}"] = np.asarray([pr, rc, f1, r.mapk(), r.mean_ndcg(), r.mrr()], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "ranking.npz"), **out)
    print("ranking:", {k: out[k] for k in out if k.startswith("k")})

    # ---- classification metrics
    yt = (torch.rand(500, 1, generator=g) < 0.35).float()
    yp = torch.rand(500, 1, generator=g)
    yp[::17] = 0.5                                         # exactly on the threshold: counts as positive
    ev = {"y_true": yt.numpy(), "y_pred": yp.numpy(), "metrics": np.asarray(Evaluator.eval(yt, yp), dtype=np.float64)}
    np.savez_compressed(os.path.join(OUT, "evaluator.npz"), **ev)
    print("evaluator:", ev["metrics"])

    # ---- feature-matrix assembly: the reference's own feature() on stand-in side tables (no dataset files needed)
    import pandas as pd
    from data.reader import MovieLens100K
    nu, ni, B = 30, 40, 500
    fu = feature_matrix(g, nu, nu, ni)[:, 2:26].numpy().astype(np.float64)          # age, gender(2), occupation(21)
    fi = feature_matrix(g, ni, nu, ni)[:, 26:].numpy().astype(np.float64)           # genre(19)
    user_data = pd.DataFrame(fu, columns=["age"] + [f"u{k}" for k in range(23)])
    user_data.insert(0, "user_id", np.arange(nu))
    item_data = pd.DataFrame(fi, columns=[f"m{k}" for k in range(19)])
    item_data.insert(0, "item_id", np.arange(ni))
    pairs = pd.DataFrame({"user_id": torch.randint(0, nu, (B,), generator=g).numpy(),
                          "item_id": torch.randint(0, ni, (B,), generator=g).numpy(),
                          "rating": (torch.rand(B, generator=g) < 0.5).long().numpy()})

    class Stub:
        pass
    stub = Stub()
    stub.user_data, stub.item_data = user_data, item_data
    feat = MovieLens100K.feature(stub, pairs)
    rating = feat.iloc[:, 2].values.astype(np.float32)                               # scripts/deepfm.py:42-44
    feat = feat.drop("rating", axis=1)
    np.savez_compressed(os.path.join(OUT, "features.npz"), users=pairs["user_id"].values, items=pairs["item_id"].values,
                        pair_rating=pairs["rating"].values.astype(np.float32), user_feat=fu.astype(np.float32),
                        item_feat=fi.astype(np.float32), x=feat.values.astype(np.float32), rating=rating)
    print("features:", feat.shape)


if __name__ == "__main__":
    main()
