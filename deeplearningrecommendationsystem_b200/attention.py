"""Attention pooling (AFM), target attention (DIN / DIEN) and the GRU recurrence of the drop-in modules."""
import torch


def afm_pool(E, W, b, h):
    """Attention pooling over the pairwise Hadamard products of E (B, F, D)            reference model/afm.py:55-65."""
    F = E.shape[1]
    iu = torch.triu_indices(F, F, offset=1, device=E.device)
    P = E[:, iu[0]] * E[:, iu[1]]
    a = torch.relu(torch.matmul(P, W) + b)
    w = torch.softmax(torch.matmul(a, h), dim=1)
    return (w * P).sum(dim=1)


def din_attention(hist_embed, target_embed, unit, pool):
    """softmax_L(MLP([h, h-t, t])) applied to the history                               reference model/din.py:39-47.
    pool=True -> (B, D) weighted sum; pool=False -> (B, L, D) scaled history (model/dien.py:33-37)."""
    t = target_embed.unsqueeze(1).expand_as(hist_embed)
    w = torch.softmax(unit(torch.cat([hist_embed, hist_embed - t, t], dim=-1)).squeeze(-1), dim=-1)
    scaled = hist_embed * w.unsqueeze(-1)
    return scaled.sum(dim=1) if pool else scaled


def gru_last_hidden(x, gru):
    """hidden[-1] of a single-layer batch_first GRU started from zeros                  reference model/dien.py:61-64."""
    _, hidden = gru(x)
    return hidden[-1]
