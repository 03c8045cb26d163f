"""Matrix factorisation sigmoid(<U[u], V[i]>) -- drop-in for reference model/mf.py:11-35 (note the 1-D output)."""
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from . import _blocks as K
from .. import ops


class MatrixFactorization(nn.Module, K.FusedRows):
    def __init__(self, num_users: int, num_items: int, embedding_size: int):
        super().__init__()
        self.user_embeddings = nn.Embedding(num_users, embedding_size)
        self.item_embeddings = nn.Embedding(num_items, embedding_size)
        xavier_normal_(self.user_embeddings.weight.data)
        xavier_normal_(self.item_embeddings.weight.data)

    def _fused_groups(self):
        return [[self.user_embeddings, self.item_embeddings]]

    def forward(self, user_indices: torch.Tensor, item_indices: torch.Tensor) -> torch.Tensor:
        dot = K.pair_lookup(self.user_embeddings.weight, self.item_embeddings.weight, user_indices, item_indices, "dot2")
        return torch.sigmoid(dot)                                   # (B,)

    def recommendation(self, num_users, num_items):
        # U @ V^T + topk (reference model/mf.py:28-35) fused: one CTA per user scores and ranks the catalogue in shared memory
        with torch.no_grad():
            return ops.mf_rank(self.user_embeddings.weight[:num_users], self.item_embeddings.weight[:num_items], num_items).cpu().numpy()
