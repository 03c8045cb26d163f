"""CPU restatement of one train step of the synthetic N-field FM / FFM configs (BASELINE.json configs[1]).

The reference modules are hard-wired to the six MovieLens features, so for F = 26 this file applies the
same arithmetic (oracle/interactions.py, validated against the real modules at F = 6 through
tests/golden) to F id-fields, with the reference's update semantics: nn.Embedding's DENSE gradient
(embedding_dense_backward) followed by plain SGD over every row.
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Also the timed CPU baseline ("port") in bench.py.
"""
import torch

from . import interactions as I


def gather_fields(table, ids, offsets):
    """table (R_total, W); ids (B, F) local ids; offsets (F,) row offset of each field -> (B, F, W)."""
    return table[ids + offsets.unsqueeze(0)]


def fm_logit(table, ids, offsets, bias):
    return I.fm_second_order(gather_fields(table, ids, offsets)) + bias


def ffm_logit(table, ids, offsets, bias, fast=False):
    B, F = ids.shape
    T = gather_fields(table, ids, offsets).view(B, F, F, -1)
    return (I.ffm_cross_fast(T) if fast else I.ffm_cross(T)) + bias


def train_step(kind, table, bias, ids, offsets, y, lr, fast=False):
    """fwd + BCE + bwd + dense SGD, in place on `table`/`bias`.  Returns (pred, loss)."""
    table.requires_grad_(True)
    bias.requires_grad_(True)
    logit = (fm_logit if kind == "fm" else lambda *a: ffm_logit(*a, fast=fast))(table, ids, offsets, bias)
    pred = torch.sigmoid(logit)
    loss = I.bce(pred, y)
    gt, gb = torch.autograd.grad(loss, [table, bias])
    table.requires_grad_(False)
    bias.requires_grad_(False)
    table -= lr * gt
    bias -= lr * gb
    return pred.detach(), loss.detach()


class _PortFields(torch.nn.Module):
    """F id-fields in one concatenated nn.Embedding (dense gradient, as the reference's per-feature nn.Embedding tables)."""

    def __init__(self, cards, dim):
        super().__init__()
        self.table = torch.nn.Embedding(sum(cards), dim)
        self.register_buffer("offsets", torch.tensor([sum(cards[:i]) for i in range(len(cards))]))
        for k, c in enumerate(cards):
            o = int(self.offsets[k])
            torch.nn.init.normal_(self.table.weight.data[o:o + c], 0.0, (2.0 / (c + dim)) ** 0.5)

    def fields(self, ids):
        return self.table(ids + self.offsets.unsqueeze(0))              # (B, F, D)


class PortPNN(_PortFields):
    """Inner-product PNN over F id-fields: the arithmetic of model/pnn.py:55-66,75-77,111-131 (oracle/ml100k.pnn, validated
    against the real module at F = 6) applied to F fields.  Timed CPU baseline of the C3 config."""

    def __init__(self, cards, dim, hidden):
        super().__init__(cards, dim)
        F = len(cards)
        self.linear1 = torch.nn.Linear(F * dim, hidden[0])
        self.linear2 = torch.nn.Linear(F * (F - 1) // 2, hidden[0])
        self.dnn = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in zip(hidden[:-1], hidden[1:])])
        self.output = torch.nn.Linear(hidden[-1], 1)

    def forward(self, ids):
        E = self.fields(ids)
        h = self.linear1(E.flatten(1)) + self.linear2(I.inner_products(E))
        for layer in self.dnn:
            h = torch.relu(layer(h))
        return torch.sigmoid(self.output(h)).view(-1, 1)


class PortAFM(_PortFields):
    """Attentional pooling over the F(F-1)/2 pair products of F id-fields: model/afm.py:55-66 (oracle/interactions.afm_pool,
    which materialises the (B, P, D) pair tensor exactly like the reference).  Timed CPU baseline of the C3 config."""

    def __init__(self, cards, dim, att):
        super().__init__(cards, dim)
        self.W = torch.nn.Parameter(torch.randn(dim, att) * 0.1)
        self.b = torch.nn.Parameter(torch.zeros(att))
        self.h = torch.nn.Parameter(torch.randn(att, 1) * 0.1)
        self.output_layer = torch.nn.Linear(dim, 1)

    def forward(self, ids):
        return torch.sigmoid(self.output_layer(I.afm_pool(self.fields(ids), self.W, self.b, self.h)))
