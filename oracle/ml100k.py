"""Functional CPU restatement of the ten hot-path models, keyed by the reference's state_dict names.

``forward(name, sd, *inputs)`` evaluates model ``name`` from a plain ``{key: tensor}`` dict whose keys and
shapes are those of the reference modules, so golden state_dicts load as-is and
``torch.autograd`` gives the oracle gradients.  The structure deliberately differs from the reference
(one table-driven embedding step + the N-field interaction functions) -- it is a restatement, not a copy.
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Feature-matrix layout (data/reader.py:98-101): x (B,45) f32 = [uid, iid, age, gender(2), occupation(21),
genre(19)].
"""
import torch

from . import interactions as I

# (column start, width) of each side feature inside x
AGE, GENDER, OCC, GENRE = (2, 1), (3, 2), (5, 21), (26, 19)


def _ids(x, col):
    return x[:, col].long()


def _bag(x, span, W):
    """one-/multi-hot 'lookup' done as a dense matmul (model/deepfm.py:47-51)."""
    s, w = span
    return x[:, s:s + w] @ W


def _six_fields(x, sd, names):
    """[user, item, age, gender, occupation, movie] embeddings, each (B, D)."""
    u, i, a, g, o, m = (sd[n] for n in names)
    return [u[_ids(x, 0)], i[_ids(x, 1)], _bag(x, AGE, a), _bag(x, GENDER, g), _bag(x, OCC, o),
            _bag(x, GENRE, m)]


def _first_order(x, sd, lin):
    """user(1) + item(1) + Linear(43,1)(x[:,2:])            model/lr.py:24-25."""
    return sd["user.weight"][_ids(x, 0)] + sd["item.weight"][_ids(x, 1)] + \
        x[:, 2:] @ sd[f"{lin}.weight"].t() + sd[f"{lin}.bias"]


def _layers(sd, prefix, idxs):
    return [(sd[f"{prefix}.{k}.weight"], sd[f"{prefix}.{k}.bias"]) for k in idxs]


def _count(sd, prefix):
    return sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in sd if k.startswith(prefix + ".")})


_EMB6 = ["user_embedding.weight", "item_embedding.weight", "age_embedding.weight",
         "gender_embedding.weight", "occupation_embedding.weight", "movie_embedding.weight"]


def lr(sd, x):
    return torch.sigmoid(_first_order(x, sd, "linear"))


def mf(sd, u, i):
    """model/mf.py:23-26 -- note the 1-D (B,) output."""
    return torch.sigmoid((sd["user_embeddings.weight"][u] * sd["item_embeddings.weight"][i]).sum(dim=1))


def deepfm(sd, x):
    """model/deepfm.py:43-83."""
    E = torch.stack(_six_fields(x, sd, _EMB6), dim=1)
    deep = E.flatten(1) @ sd["linear.weight"].t() + sd["linear.bias"]
    deep = I.relu_tower(deep, _layers(sd, "dnn_network", _count(sd, "dnn_network")))
    wide = _first_order(x, sd, "wide") + I.fm_second_order(E).unsqueeze(1)
    z = torch.cat([wide, deep], dim=1) @ sd["output.weight"].t() + sd["output.bias"]
    return torch.sigmoid(z)


def nfm(sd, x):
    """model/nfm.py:43-72."""
    E = torch.stack(_six_fields(x, sd, _EMB6), dim=1)
    deep = I.bi_interaction(E) @ sd["linear.weight"].t() + sd["linear.bias"]
    deep = I.relu_tower(deep, _layers(sd, "dnn_network", _count(sd, "dnn_network")))
    wide = _first_order(x, sd, "wide")
    z = torch.cat([wide, deep], dim=1) @ sd["output.weight"].t() + sd["output.bias"]
    return torch.sigmoid(z)


def afm(sd, x):
    """model/afm.py:41-71 -- age enters as a scalar broadcast over D, there is no age table."""
    u = sd["user_embedding.weight"][_ids(x, 0)]
    D = u.shape[1]
    fields = [u, sd["item_embedding.weight"][_ids(x, 1)], x[:, 2:3].expand(-1, D),
              _bag(x, GENDER, sd["gender_embedding.weight"]), _bag(x, OCC, sd["occupation_embedding.weight"]),
              _bag(x, GENRE, sd["movie_embedding.weight"])]
    E = torch.stack(fields, dim=1)
    pooled = I.afm_pool(E, sd["attention_W"], sd["attention_b"], sd["attention_h"])
    cross = pooled @ sd["output_layer.weight"].t() + sd["output_layer.bias"]
    return torch.sigmoid(_first_order(x, sd, "linear") + cross)


# feature order used for T and the field (0 = user, 1 = item) each feature belongs to (model/ffm.py:9-22)
_FFM_FEATS = ["age", "gender", "occupation", "movie", "userid", "itemid"]
_FFM_FIELD_OF = [0, 0, 0, 1, 0, 1]


def ffm(sd, x):
    """model/ffm.py:46-86, including the quirk that the scalar cross term is added to every raw feature
    before the linear layer: logit = u1 + i1 + W.(x[:,2:] + cross) + b."""
    spans = {"age": AGE, "gender": GENDER, "occupation": OCC, "movie": GENRE}
    rows = []
    for f in _FFM_FEATS:
        per_field = []
        for fld in ("user", "item"):
            W = sd[f"{f}_{fld}.weight"]
            if f == "userid":
                per_field.append(W[_ids(x, 0)])
            elif f == "itemid":
                per_field.append(W[_ids(x, 1)])
            else:
                per_field.append(_bag(x, spans[f], W))
        rows.append(torch.stack(per_field, dim=1))
    T = torch.stack(rows, dim=1)                            # (B, 6, 2, D)
    cross = I.ffm_cross(T, _FFM_FIELD_OF)
    lin = (x[:, 2:] + cross.unsqueeze(1)) @ sd["linear.weight"].t() + sd["linear.bias"]
    return torch.sigmoid(sd["user.weight"][_ids(x, 0)] + sd["item.weight"][_ids(x, 1)] + lin)


_PNN6 = ["user_embed.weight", "item_embed.weight", "age_embed.weight", "gender_embed.weight",
         "occupation_embed.weight", "movie_embed.weight"]


def pnn(sd, x, mode="in"):
    """model/pnn.py:55-77,111-131.  'out' collapses the batch (S^T S) and only broadcasts when B == D."""
    E = torch.stack(_six_fields(x, sd, _PNN6), dim=1)
    lz = E.flatten(1) @ sd["product.linear1.weight"].t() + sd["product.linear1.bias"]
    if mode == "in":
        p = I.inner_products(E)
    elif mode == "out":
        p = I.outer_product_pooled(E)
        if p.shape[0] != E.shape[0]:
            raise RuntimeError("PNN 'out': (D,H0) product term cannot broadcast against (B,H0) unless B == D")
    else:
        raise ValueError(mode)
    lp = p @ sd["product.linear2.weight"].t() + sd["product.linear2.bias"]
    h = I.relu_tower(lz + lp, _layers(sd, "dnn.dnn_network", _count(sd, "dnn.dnn_network")))
    return torch.sigmoid(h @ sd["output.weight"].t() + sd["output.bias"]).view(-1, 1)


def _din_core(sd, hist, tgt, pfx):
    tab = sd[f"{pfx}item_embedding.weight"]
    t, h = tab[tgt], tab[hist]
    w = I.din_attention(h, t, _layers(sd, f"{pfx}attention", (0, 2, 4)))
    return h, t, w


def din(sd, hist, tgt):
    """model/din.py:33-53."""
    h, t, w = _din_core(sd, hist, tgt, "")
    pooled = (h * w.unsqueeze(-1)).sum(dim=1)
    z = I.relu_tower(torch.cat([pooled, t], dim=1), _layers(sd, "fc", (0, 2, 4)), relu_last=False)
    return torch.sigmoid(z)


def dien(sd, hist, tgt):
    """model/dien.py:23-39,57-68 -- a plain GRU over the attention-scaled history (not AUGRU)."""
    h, t, w = _din_core(sd, hist, tgt, "din.")
    hL = I.gru(h * w.unsqueeze(-1), sd["interest_evolution.weight_ih_l0"], sd["interest_evolution.weight_hh_l0"],
               sd["interest_evolution.bias_ih_l0"], sd["interest_evolution.bias_hh_l0"])
    z = I.relu_tower(torch.cat([hL, t], dim=-1), _layers(sd, "fc", (0, 2, 4)), relu_last=False)
    return torch.sigmoid(z)


def neuralcf(sd, u, i):
    """model/neuralcf.py:33-59."""
    gmf = sd["GMF_Embedding_User.weight"][u] * sd["GMF_Embedding_Item.weight"][i]
    z = torch.cat([sd["MLP_Embedding_User.weight"][u], sd["MLP_Embedding_Item.weight"][i]], dim=1)
    z = I.relu_tower(z, _layers(sd, "dnn_network", _count(sd, "dnn_network")))
    mlp = z @ sd["linear.weight"].t() + sd["linear.bias"]
    return torch.sigmoid(torch.cat([gmf, mlp], dim=1) @ sd["linear2.weight"].t() + sd["linear2.bias"])


def _stacked(x, sd):
    """[e_user, e_item, age, e_gender, e_occupation, e_movie] (B, 5D+1)   model/widedeep.py:44-50."""
    return torch.cat([sd["user_embedding.weight"][_ids(x, 0)], sd["item_embedding.weight"][_ids(x, 1)], x[:, 2:3],
                      _bag(x, GENDER, sd["gender_embedding.weight"]), _bag(x, OCC, sd["occupation_embedding.weight"]),
                      _bag(x, GENRE, sd["movie_embedding.weight"])], dim=1)


def widedeep(sd, x):
    """model/widedeep.py:42-61 -- no ReLU after the first projection."""
    deep = _stacked(x, sd) @ sd["linear.weight"].t() + sd["linear.bias"]
    deep = I.relu_tower(deep, _layers(sd, "dnn_network", _count(sd, "dnn_network")))
    z = torch.cat([_first_order(x, sd, "wide"), deep], dim=1) @ sd["output.weight"].t() + sd["output.bias"]
    return torch.sigmoid(z)


def deepcross(sd, x):
    """model/deepcross.py:10-18 (x_{l+1} = x_0 * (W_l x_l) + b_l + x_l, full-matrix W) and :58-72."""
    z0 = _stacked(x, sd)
    c = z0
    for l in _count(sd, "cross_network.cross_weights"):
        c = z0 * (c @ sd[f"cross_network.cross_weights.{l}.weight"].t()) + sd[f"cross_network.cross_biases.{l}"] + c
    deep = I.relu_tower(z0, _layers(sd, "deep_network.network", _count(sd, "deep_network.network")))
    both = torch.cat([c, deep], dim=1)
    return torch.sigmoid(both @ sd["output_layer.weight"].t() + sd["output_layer.bias"])


def deepcrossing(sd, x):
    """model/deepcrossing.py:20-25 residual units and :54-69."""
    r = _stacked(x, sd)
    for l in _count(sd, "res_layers"):
        h = torch.relu(r @ sd[f"res_layers.{l}.linear1.weight"].t() + sd[f"res_layers.{l}.linear1.bias"])
        r = torch.relu(h @ sd[f"res_layers.{l}.linear2.weight"].t() + sd[f"res_layers.{l}.linear2.bias"] + r)
    return torch.sigmoid(r @ sd["linear.weight"].t() + sd["linear.bias"])


MODELS = {
    "lr": lr, "mf": mf, "deepfm": deepfm, "nfm": nfm, "afm": afm, "ffm": ffm,
    "pnn_in": lambda sd, x: pnn(sd, x, "in"), "pnn_out": lambda sd, x: pnn(sd, x, "out"),
    "din": din, "dien": dien, "neuralcf": neuralcf,
    "widedeep": widedeep, "deepcross": deepcross, "deepcrossing": deepcrossing,
}


def forward(name, sd, *inputs):
    return MODELS[name](sd, *inputs)


def loss_and_grads(name, sd, inputs, rating):
    """pred, BCE loss and d loss / d every entry of sd (dense, like the reference's nn.Embedding grads)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    pred = forward(name, leaves, *inputs)
    loss = I.bce(pred, rating)
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    out = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(leaves, grads)}
    return pred.detach(), loss.detach(), out
