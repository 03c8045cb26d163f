"""Wide & Deep -- drop-in for reference model/widedeep.py:7-79.

Deep side: five embeddings + raw age stacked to 5D+1 -> Linear -> ReLU tower (no ReLU after the first projection,
model/widedeep.py:51-54); wide side: the LR first-order term; Linear(2, 1) over [wide, deep] -> sigmoid."""
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from . import _blocks as K


class WideDeep(nn.Module):
    def __init__(self, num_users, num_items, hidden_units, embedding_dim):
        super().__init__()
        self.user_embedding = nn.Embedding(num_users, embedding_dim)
        self.item_embedding = nn.Embedding(num_items, embedding_dim)
        self.gender_embedding = nn.Embedding(2, embedding_dim)
        self.occupation_embedding = nn.Embedding(21, embedding_dim)
        self.movie_embedding = nn.Embedding(19, embedding_dim)
        self.linear = nn.Linear(embedding_dim * 5 + 1, hidden_units[0])
        self.dnn_network = nn.ModuleList([nn.Linear(a, b) for a, b in zip(hidden_units[:-1], hidden_units[1:])])
        self.relu = nn.ReLU()
        self.user = nn.Embedding(num_users, 1)
        self.item = nn.Embedding(num_items, 1)
        self.wide = nn.Linear(1 + 2 + 21 + 19, 1)
        self.output = nn.Linear(2, 1)
        for emb in (self.user_embedding, self.item_embedding, self.gender_embedding, self.occupation_embedding,
                    self.movie_embedding, self.user, self.item):
            xavier_normal_(emb.weight.data)

    def forward(self, x):
        deep = self.linear(K.stacked_features(x, self.user_embedding.weight, self.item_embedding.weight,
                                              self.gender_embedding.weight, self.occupation_embedding.weight,
                                              self.movie_embedding.weight))
        for layer in self.dnn_network:
            deep = self.relu(layer(deep))
        wide = K.first_order(self.user, self.item, self.wide, x)
        return torch.sigmoid(self.output(torch.cat((wide, deep), dim=1)))

    def recommendation(self, num_users, user_item, k):
        return K.topk_per_user(self, num_users, user_item, k)
