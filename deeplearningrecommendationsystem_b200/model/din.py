"""Deep Interest Network -- drop-in for reference model/din.py:9-66.

One item table, looked up for the target and for the L history slots in ONE gather (so the backward needs one sort
/ segment-reduce); attention unit MLP 3D->128->64->1 over [h, h-t, t]; softmax over L with no padding mask and no
scaling (padded slots hold the real item id 0, scripts/din.py:31); weighted sum; fc 2D->256->128->1->Sigmoid."""
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from . import _blocks as K
from .. import attention


class DIN(nn.Module, K.FusedRows):
    def __init__(self, num_items, embed_size):
        super().__init__()
        self.item_embedding = nn.Embedding(num_items, embed_size)
        self.attention = nn.Sequential(nn.Linear(embed_size * 3, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 1))
        self.fc = nn.Sequential(nn.Linear(embed_size * 2, 256), nn.ReLU(), nn.Linear(256, 128), nn.ReLU(), nn.Linear(128, 1),
                                nn.Sigmoid())
        xavier_normal_(self.item_embedding.weight.data)

    def _fused_groups(self):
        return [[self.item_embedding]]

    def forward(self, hist, target_item):
        rows = K.lookup(self.item_embedding.weight, torch.cat([hist, target_item.unsqueeze(1)], dim=1))   # (B, L+1, D)
        pooled = attention.din_attention(rows, self.attention, pool=True)                                # (B, D)
        return self.fc(torch.cat([pooled, rows[:, -1]], dim=1))

    def recommendation(self, num_users, num_items, hist_list, k):
        device = next(self.parameters()).device
        return K.rank_catalogue(K.history_scores(self, hist_list, num_items, device), num_users, num_items, k)
