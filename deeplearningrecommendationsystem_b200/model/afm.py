"""Attentional FM -- drop-in for reference model/afm.py:7-83.

Five embedding tables plus age as a scalar broadcast over D (there is no age table, model/afm.py:45,54); the 15
pairwise Hadamard products are attention-pooled (relu(P.W + b).h, softmax over the pairs) and projected to a logit
that is added to the first-order term."""
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from . import _blocks as K
from .. import attention


class AFM(nn.Module):
    def __init__(self, num_users, num_items, embedding_dim, attention_dim):
        super().__init__()
        self.user_embedding = nn.Embedding(num_users, embedding_dim)
        self.item_embedding = nn.Embedding(num_items, embedding_dim)
        self.gender_embedding = nn.Embedding(2, embedding_dim)
        self.occupation_embedding = nn.Embedding(21, embedding_dim)
        self.movie_embedding = nn.Embedding(19, embedding_dim)
        self.attention_W = nn.Parameter(torch.randn(embedding_dim, attention_dim))
        self.attention_b = nn.Parameter(torch.randn(attention_dim))
        self.attention_h = nn.Parameter(torch.randn(attention_dim, 1))
        self.output_layer = nn.Linear(embedding_dim, 1)
        self.user = nn.Embedding(num_users, 1)
        self.item = nn.Embedding(num_items, 1)
        self.linear = nn.Linear(1 + 2 + 21 + 19, 1)
        for emb in (self.user_embedding, self.item_embedding, self.gender_embedding, self.occupation_embedding,
                    self.movie_embedding, self.user, self.item):
            xavier_normal_(emb.weight.data)

    _SLOTS = ((K.COL_USER, 1, K.KIND_ID), (K.COL_ITEM, 1, K.KIND_ID), (K.AGE[0], 1, K.KIND_SCALAR),
              (K.GENDER[0], K.GENDER[1], K.KIND_BAG), (K.OCC[0], K.OCC[1], K.KIND_BAG), (K.GENRE[0], K.GENRE[1], K.KIND_BAG))

    def forward(self, x):
        E = K.XEmbed.apply(x, self._SLOTS, self.user_embedding.weight, self.item_embedding.weight, self.gender_embedding.weight,
                           self.occupation_embedding.weight, self.movie_embedding.weight)       # (B, 6, D), slot 2 = age
        pooled = attention.afm_pool(E, self.attention_W, self.attention_b, self.attention_h)    # (B, D)
        cross = self.output_layer(pooled)
        return torch.sigmoid(K.first_order(self.user, self.item, self.linear, x) + cross)

    def recommendation(self, num_users, user_item, k):
        return K.topk_per_user(self, num_users, user_item, k)
