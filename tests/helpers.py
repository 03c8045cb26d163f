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
