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
