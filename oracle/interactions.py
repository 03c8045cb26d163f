"""N-field restatement of the reference's feature-interaction arithmetic (fp32, torch CPU).

Every function takes the per-sample field embeddings as one tensor ``E`` of shape (B, F, D) (or, for
FFM, ``T`` of shape (B, F, NF, D)) instead of the reference's hard-wired six python variables, so the
same code is the oracle for the MovieLens modules (F = 6) and for the synthetic configs (F = 26 / 39).
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
import torch


def pair_index(F):
    """(i, j) index vectors of the i<j pairs in the reference's nested-loop order
    (model/afm.py:57-59, model/nfm.py:60-62, model/pnn.py:61-65)."""
    iu = torch.triu_indices(F, F, offset=1)
    return iu[0], iu[1]


def fm_second_order(E):
    """0.5 * sum_d[(sum_f e)^2 - sum_f e^2]  -> (B,)      model/deepfm.py:71-76."""
    s = E.sum(dim=1)
    return 0.5 * (s * s - (E * E).sum(dim=1)).sum(dim=1)


def bi_interaction(E):
    """sum_{i<j} e_i * e_j  -> (B, D), accumulated pair by pair as model/nfm.py:59-62 does."""
    F = E.shape[1]
    acc = torch.zeros_like(E[:, 0])
    for i in range(F):
        for j in range(i + 1, F):
            acc = acc + E[:, i] * E[:, j]
    return acc


def inner_products(E):
    """[<e_i, e_j>]_{i<j} -> (B, F(F-1)/2)                  model/pnn.py:61-66."""
    i, j = pair_index(E.shape[1])
    return (E[:, i] * E[:, j]).sum(dim=2)


def outer_product_pooled(E):
    """S = sum_f e_f ; S^T S -> (D, D), collapsed over the batch      model/pnn.py:69-72."""
    s = E.sum(dim=1)
    return s.t() @ s


def afm_pool(E, W, b, h):
    """Attention pooling over the pairwise Hadamard products            model/afm.py:55-65.
    E (B,F,D), W (D,A), b (A,), h (A,1) -> (B, D)."""
    i, j = pair_index(E.shape[1])
    P = E[:, i] * E[:, j]                                   # (B, P, D)
    a = torch.relu(P @ W + b)                               # (B, P, A)
    w = torch.softmax(a @ h, dim=1)                         # (B, P, 1)
    return (w * P).sum(dim=1)


def ffm_cross(T, field_of=None):
    """sum_{i<j} <v_{i,field(j)}, v_{j,field(i)}>  -> (B,)              model/ffm.py:61-82.
    T (B, F, NF, D): T[:, i, c] is feature i's vector toward field c.  ``field_of[j]`` is the field
    feature j belongs to (identity when every feature is its own field, NF == F)."""
    F = T.shape[1]
    if field_of is None:
        field_of = list(range(F))
    out = torch.zeros(T.shape[0], dtype=T.dtype)
    for i in range(F):
        for j in range(i + 1, F):
            out = out + (T[:, i, field_of[j]] * T[:, j, field_of[i]]).sum(dim=1)
    return out


def ffm_cross_fast(T):
    """Same value as ffm_cross for NF == F, vectorised (used only as the timed CPU baseline)."""
    F = T.shape[1]
    i, j = pair_index(F)
    return (T[:, i, j] * T[:, j, i]).sum(dim=(1, 2))


def din_attention(h, t, att):
    """Target attention weights over a behaviour sequence               model/din.py:39-44.
    h (B,L,D), t (B,D), att = [(W0,b0),(W1,b1),(W2,b2)] -> softmax weights (B, L) (no mask)."""
    te = t.unsqueeze(1).expand_as(h)
    z = torch.cat([h, h - te, te], dim=-1)
    (W0, b0), (W1, b1), (W2, b2) = att
    z = torch.relu(z @ W0.t() + b0)
    z = torch.relu(z @ W1.t() + b1)
    z = (z @ W2.t() + b2).squeeze(-1)
    return torch.softmax(z, dim=-1)


def gru(x, w_ih, w_hh, b_ih, b_hh):
    """Single-layer batch_first GRU with h0 = 0, gate order (r, z, n)   model/dien.py:47,61 (nn.GRU).
    x (B,L,D) -> final hidden (B, H)."""
    B, L, _ = x.shape
    H = w_hh.shape[1]
    h = torch.zeros(B, H, dtype=x.dtype)
    for s in range(L):
        gi = x[:, s] @ w_ih.t() + b_ih
        gh = h @ w_hh.t() + b_hh
        ir, iz, inn = gi.split(H, dim=1)
        hr, hz, hn = gh.split(H, dim=1)
        r = torch.sigmoid(ir + hr)
        z = torch.sigmoid(iz + hz)
        n = torch.tanh(inn + r * hn)
        h = (1.0 - z) * n + z * h
    return h


def relu_tower(x, layers, relu_last=True):
    """Linear+ReLU stack; the reference applies ReLU after EVERY layer incl. the last
    (model/deepfm.py:58-60, model/pnn.py:19-21, model/neuralcf.py:48-50)."""
    for k, (W, b) in enumerate(layers):
        x = x @ W.t() + b
        if relu_last or k + 1 < len(layers):
            x = torch.relu(x)
    return x


def bce(p, y):
    """mean(-[y*max(log p,-100) + (1-y)*max(log(1-p),-100)])   torch.nn.BCELoss (scripts/deepfm.py:54)."""
    lp = torch.clamp(torch.log(p), min=-100.0)
    l1p = torch.clamp(torch.log(1.0 - p), min=-100.0)
    return -(y * lp + (1.0 - y) * l1p).mean()
