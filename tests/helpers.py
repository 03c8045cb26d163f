"""Shared seeded input builders for tests and the golden generator."""
import torch


def feature_matrix(g, B, num_users, num_items):
    """(B,45) float32 in the data/reader.py:98-101 column order."""
    x = torch.zeros(B, 45)
    x[:, 0] = torch.randint(0, num_users, (B,), generator=g).float()
    x[:, 1] = torch.randint(0, num_items, (B,), generator=g).float()
    x[:, 2] = torch.rand(B, generator=g)
    gender = torch.randint(0, 2, (B,), generator=g)
    x[torch.arange(B), 3 + gender] = 1.0
    occ = torch.randint(0, 21, (B,), generator=g)
    x[torch.arange(B), 5 + occ] = 1.0
    n_genre = torch.randint(0, 7, (B,), generator=g)  # 0..6 active genres (reader: max 6)
    for b in range(B):
        idx = torch.randperm(19, generator=g)[: int(n_genre[b])]
        x[b, 26 + idx] = 1.0
    return x


def catalogue_frame(g, num_users, num_items, shuffle=True):
    """Every (user, item) pair with that user's and item's side features, as the pandas frame `data.user_item()`
    builds (data/reader.py:104-112): columns [user_id, item_id, age, gender(2), occupation(21), genre(19)].
    Rows are shuffled so that grouping by user has real work to do; each user keeps num_items rows."""
    import numpy as np
    import pandas as pd
    fu = feature_matrix(g, num_users, num_users, num_items)      # per-user side features (cols 2..25)
    fi = feature_matrix(g, num_items, num_users, num_items)      # per-item genres (cols 26..44)
    u = torch.arange(num_users).repeat_interleave(num_items)
    i = torch.arange(num_items).repeat(num_users)
    x = torch.zeros(num_users * num_items, 45)
    x[:, 0], x[:, 1] = u.float(), i.float()
    x[:, 2:26] = fu[u, 2:26]
    x[:, 26:] = fi[i, 26:]
    if shuffle:
        x = x[torch.randperm(x.shape[0], generator=g)]
    cols = ["user_id", "item_id", "age"] + [f"g{k}" for k in range(2)] + [f"o{k}" for k in range(21)] + [f"m{k}" for k in range(19)]
    df = pd.DataFrame(x.numpy().astype(np.float64), columns=cols)
    df["user_id"] = df["user_id"].astype(np.int64)
    df["item_id"] = df["item_id"].astype(np.int64)
    return df


def feature_matrix_fast(g, B, nu, ni):
    """vectorised feature_matrix for large B (same column layout; n distinct random genres per row)"""
    x = torch.zeros(B, 45)
    x[:, 0] = torch.randint(0, nu, (B,), generator=g).float()
    x[:, 1] = torch.randint(0, ni, (B,), generator=g).float()
    x[:, 2] = torch.rand(B, generator=g)
    r = torch.arange(B)
    x[r, 3 + torch.randint(0, 2, (B,), generator=g)] = 1.0
    x[r, 5 + torch.randint(0, 21, (B,), generator=g)] = 1.0
    n_genre = torch.randint(0, 7, (B, 1), generator=g)
    order = torch.rand(B, 19, generator=g).argsort(dim=1)
    x[:, 26:] = (order < n_genre).float()
    return x
