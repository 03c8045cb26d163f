#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

For every hot-path model it instantiates the reference nn.Module from
/root/reference/model/*.py under a fixed seed, feeds seeded inputs, and records

  in/<name>      the forward inputs
  sd/<key>       the initial state_dict
  pred, loss     forward output and BCELoss (reference trainer/trainer.py:33-37)
  grad/<key>     every parameter gradient after loss.backward() (trainer/trainer.py:38)
  sd2/<key>      the state_dict after TWO Trainer.train_loop steps with the script optimiser
                 Adam(lr=1e-3, weight_decay=1e-5) (scripts/deepfm.py:55, trainer/trainer.py:23-40)
  losses         the two training losses

plus the sampler's id streams (sampler/sampler.py:16-48) under a fixed `random.seed`.
Nothing from the reference is copied into the repo; only these numeric vectors are committed.
"""
import os
import random
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(OUT))
from helpers import feature_matrix  # noqa: E402

from model.lr import LogisticRegression  # noqa: E402
from model.mf import MatrixFactorization  # noqa: E402
from model.ffm import FFM  # noqa: E402
from model.deepfm import DeepFM  # noqa: E402
from model.afm import AFM  # noqa: E402
from model.nfm import NFM  # noqa: E402
from model.pnn import PNN  # noqa: E402
from model.din import DIN  # noqa: E402
from model.dien import DIEN  # noqa: E402
from model.neuralcf import NeuralCF  # noqa: E402
from trainer.trainer import Trainer  # noqa: E402
from sampler.sampler import Sampler  # noqa: E402


def run(name, ctor, inputs, rating):
    torch.manual_seed(7)
    model = ctor()
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    rec = {}
    for i, t in enumerate(inputs):
        rec[f"in/{i}"] = t.numpy()
    rec["rating"] = rating.numpy()
    for k, v in sd0.items():
        rec[f"sd/{k}"] = v.numpy()
    loss_fn = torch.nn.BCELoss()
    model.zero_grad()
    pred = model(*inputs)
    loss = loss_fn(pred, rating)
    loss.backward()
    rec["pred"] = pred.detach().numpy()
    rec["loss"] = np.float32(loss.item())
    for k, p in model.named_parameters():
        rec[f"grad/{k}"] = p.grad.detach().numpy().copy()
    # two steps of the reference trainer with the script optimiser
    model.load_state_dict(sd0)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    tr = Trainer(model, loss_fn, opt)
    losses = []
    for _ in range(2):
        tr.train_loop(*inputs, train_rating=rating)
        losses.append(tr.train_loss.item())
    rec["losses"] = np.asarray(losses, dtype=np.float32)
    for k, v in model.state_dict().items():
        rec[f"sd2/{k}"] = v.numpy()
    tr.valid_loop(*inputs, valid_rating=rating)
    rec["pred_after"] = tr.predictions_valid.numpy()
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **rec)
    print(f"{name}: pred{tuple(pred.shape)} loss={loss.item():.6f} keys={len(rec)}")


def main():
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(2024)
    B, NU, NI = 96, 50, 60
    x = feature_matrix(g, B, NU, NI)
    y = (torch.rand(B, 1, generator=g) < 0.4).float()
    run("lr", lambda: LogisticRegression(NU, NI, 43), [x], y)
    run("deepfm", lambda: DeepFM(NU, NI, [32, 16, 8, 1], 16), [x], y)
    run("nfm", lambda: NFM(NU, NI, [32, 16, 8, 1], 16), [x], y)
    run("afm", lambda: AFM(NU, NI, 16, 8), [x], y)
    # FFM and PNN hard-code 943/1682 rows (model/ffm.py:19-22, model/pnn.py:87-88)
    xb = feature_matrix(g, B, 943, 1682)
    run("ffm", lambda: FFM(43, 8), [xb], y)
    run("pnn_in", lambda: PNN(8, [32, 16, 8, 4], "in"), [xb], y)
    # "out" mode only broadcasts when B == embed_dim
    xo = feature_matrix(g, 16, 943, 1682)
    yo = (torch.rand(16, 1, generator=g) < 0.4).float()
    run("pnn_out", lambda: PNN(16, [32, 16, 8, 4], "out"), [xo], yo)
    u = torch.randint(0, NU, (B,), generator=g)
    it = torch.randint(0, NI, (B,), generator=g)
    run("mf", lambda: MatrixFactorization(NU, NI, 16), [u, it], y[:, 0].clone())
    run("neuralcf", lambda: NeuralCF(NU, NI, 8, [32, 16, 8]), [u, it], y)
    hist = torch.randint(0, NI, (B, 7), generator=g)
    hist[:, :2] = 0  # left padding with the real item id 0 (scripts/din.py:31)
    tgt = torch.randint(0, NI, (B,), generator=g)
    run("din", lambda: DIN(NI, 16), [hist, tgt], y)
    run("dien", lambda: DIEN(NI, 16), [hist, tgt], y)

    # sampler streams (python `random`, sampler/sampler.py:21-27)
    rnd = random.Random(5)
    excl = {(rnd.randrange(20), rnd.randrange(30)) for _ in range(150)}
    random.seed(123)
    s = Sampler()
    u1, i1, r1 = s.negative_sampling(20, 30, excl, 5)
    u2, i2, r2 = s.negative_sampling(20, 30, excl, 2)  # state accumulates on the instance
    random.seed(321)
    df = Sampler().negative_sampling2(20, 30, excl, 3)
    np.savez_compressed(
        os.path.join(OUT, "sampler.npz"),
        excl=np.asarray(sorted(excl), dtype=np.int64),
        u1=u1.numpy(), i1=i1.numpy(), r1=r1.numpy(),
        u2=u2.numpy(), i2=i2.numpy(), r2=r2.numpy(),
        df_user=df["user_id"].values, df_item=df["item_id"].values, df_rating=df["rating"].values,
    )
    print("sampler:", len(u1), len(u2), len(df))


if __name__ == "__main__":
    main()
