"""Integer / index work restated with numpy and python `random`.  TEST INFRASTRUCTURE ONLY.

Bit-exact obligations of the hot path: gather is a pure row copy, dedup must equal
torch.unique(sorted=True, return_inverse=True, return_counts=True), and negative sampling must replay
sampler/sampler.py:21-27's python-`random` stream.
"""
import random

import numpy as np


def gather(table, ids):
    """rows of `table` (numpy, (R, W)) selected by ids -- nn.Embedding forward (model/deepfm.py:45-46)."""
    return np.take(table, np.asarray(ids).reshape(-1), axis=0).reshape(*np.shape(ids), table.shape[1])


def dedup(ids):
    """(unique ascending, inverse, counts) of a flat int64 array."""
    u, inv, cnt = np.unique(np.asarray(ids).reshape(-1), return_inverse=True, return_counts=True)
    return u.astype(np.int64), inv.reshape(-1).astype(np.int64), cnt.astype(np.int64)


def stable_order(ids):
    """positions sorted by (id, position): the deterministic segment order the update kernels use."""
    return np.argsort(np.asarray(ids).reshape(-1), kind="stable").astype(np.int64)


def negative_draws(num_user, num_item, excluded, num_negatives, rng=random):
    """The (user, item) stream of sampler/sampler.py:21-27: for each user in order, `num_negatives`
    accepted draws of rng.randint(0, num_item-1), redrawing while the pair is excluded."""
    users, items = [], []
    for u in range(num_user):
        need = num_negatives
        while need:
            it = rng.randint(0, num_item - 1)
            if (u, it) not in excluded:
                users.append(u)
                items.append(it)
                need -= 1
    return users, items
