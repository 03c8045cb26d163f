"""Deep Crossing -- drop-in for reference model/deepcrossing.py:7-85.

Stacked 5D+1 input through residual units relu(W2 relu(W1 x + b1) + b2 + x) (model/deepcrossing.py:20-25), then
Linear(5D+1, 1) -> sigmoid.  Constructor argument names follow the reference (num_user, num_item, num_feature)."""
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from . import _blocks as K


class ResidualBlock(nn.Module):
    def __init__(self, hidden_unit, dim_stack):
        super().__init__()
        self.linear1 = nn.Linear(dim_stack, hidden_unit)
        self.linear2 = nn.Linear(hidden_unit, dim_stack)
        self.relu = nn.ReLU()

    def forward(self, x):
        return self.relu(self.linear2(self.relu(self.linear1(x))) + x)


class DeepCrossing(nn.Module):
    def __init__(self, num_user, num_item, num_feature, hidden_units):
        super().__init__()
        self.user_embedding = nn.Embedding(num_user, num_feature)
        self.item_embedding = nn.Embedding(num_item, num_feature)
        self.gender_embedding = nn.Embedding(2, num_feature)
        self.occupation_embedding = nn.Embedding(21, num_feature)
        self.movie_embedding = nn.Embedding(19, num_feature)
        for emb in (self.user_embedding, self.item_embedding, self.gender_embedding, self.occupation_embedding,
                    self.movie_embedding):
            xavier_normal_(emb.weight.data)
        dim_stack = num_feature * 5 + 1
        self.res_layers = nn.ModuleList([ResidualBlock(unit, dim_stack) for unit in hidden_units])
        self.linear = nn.Linear(dim_stack, 1)

    def forward(self, feature_vector):
        r = K.stacked_features(feature_vector, self.user_embedding.weight, self.item_embedding.weight,
                               self.gender_embedding.weight, self.occupation_embedding.weight, self.movie_embedding.weight)
        for res in self.res_layers:
            r = res(r)
        return torch.sigmoid(self.linear(r))

    def recommendation(self, num_users, user_item, k):
        return K.topk_per_user(self, num_users, user_item, k)
