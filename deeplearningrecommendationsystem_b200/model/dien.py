"""DIEN as the reference defines it (model/dien.py:8-81): a DIN-style attention with a smaller unit (3D->64->32->1)
that SCALES the history instead of pooling it, an nn.GRU(D, D) over the scaled sequence (h0 = 0), and
fc 2D->128->64->1->Sigmoid on [h_L, target].  This is a plain GRU on attention-scaled inputs, not the paper's AUGRU.
state_dict keys: din.item_embedding.weight, din.attention.{0,2,4}.*, interest_evolution.*_l0, fc.{0,2,4}.*"""
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from . import _blocks as K
from .. import attention


class DIN(nn.Module, K.FusedRows):
    """Attention front half of DIEN: returns (history * attention weight (B, L, D), target embedding (B, D))."""

    def __init__(self, num_items, embed_size):
        super().__init__()
        self.item_embedding = nn.Embedding(num_items, embed_size)
        self.attention = nn.Sequential(nn.Linear(embed_size * 3, 64), nn.ReLU(), nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, 1))
        xavier_normal_(self.item_embedding.weight.data)

    def _fused_groups(self):
        return [[self.item_embedding]]

    def forward(self, hist, target_item):
        rows = K.lookup(self.item_embedding.weight, torch.cat([hist, target_item.unsqueeze(1)], dim=1))
        return attention.din_attention(rows, self.attention, pool=False), rows[:, -1]


class DIEN(nn.Module):
    def __init__(self, num_items, embed_size):
        super().__init__()
        self.din = DIN(num_items, embed_size)
        self.interest_evolution = nn.GRU(embed_size, embed_size, batch_first=True)
        self.fc = nn.Sequential(nn.Linear(embed_size * 2, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 1),
                                nn.Sigmoid())

    def fuse_embedding_updates(self):
        """opt-in fused sparse rows for the item table (see model/_blocks.RowSink); FusedRowOptimizer finds self.din"""
        self.din.fuse_embedding_updates()
        return self

    def forward(self, hist, target_item):
        att_hist, target_embed = self.din(hist, target_item)
        final_interest = attention.gru_last_hidden(att_hist, self.interest_evolution)                  # (B, D)
        return self.fc(torch.cat([final_interest, target_embed], dim=-1))

    def recommendation(self, num_users, num_items, hist_list, k):
        device = next(self.parameters()).device
        return K.rank_catalogue(K.history_scores(self, hist_list, num_items, device), num_users, num_items, k)
