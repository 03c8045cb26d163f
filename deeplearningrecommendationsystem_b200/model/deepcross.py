"""Deep & Cross -- drop-in for reference model/deepcross.py:7-89.

The reference's cross layer is the matrix form x_{l+1} = x_0 * (W_l x_l) + b_l + x_l with a full (d, d) weight
(model/deepcross.py:10-18), not the rank-one DCN-v1 layer; it runs beside a ReLU tower over the same stacked
5D+1 input and both feed one Linear -> sigmoid."""
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from . import _blocks as K


class CrossNetwork(nn.Module):
    def __init__(self, input_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        self.cross_weights = nn.ModuleList([nn.Linear(input_dim, input_dim, bias=False) for _ in range(num_layers)])
        self.cross_biases = nn.ParameterList([nn.Parameter(torch.zeros(input_dim)) for _ in range(num_layers)])

    def forward(self, x):
        x0 = x
        for w, b in zip(self.cross_weights, self.cross_biases):
            x = x0 * w(x) + b + x
        return x


class DeepNetwork(nn.Module):
    def __init__(self, input_dim, hidden_units):
        super().__init__()
        dims = [input_dim] + list(hidden_units)
        self.network = nn.Sequential(*[m for a, b in zip(dims[:-1], dims[1:]) for m in (nn.Linear(a, b), nn.ReLU())])

    def forward(self, x):
        return self.network(x)


class DeepCross(nn.Module):
    def __init__(self, num_users, num_items, cross_layers, deep_hidden_units, embedding_dim):
        super().__init__()
        self.user_embedding = nn.Embedding(num_users, embedding_dim)
        self.item_embedding = nn.Embedding(num_items, embedding_dim)
        self.gender_embedding = nn.Embedding(2, embedding_dim)
        self.occupation_embedding = nn.Embedding(21, embedding_dim)
        self.movie_embedding = nn.Embedding(19, embedding_dim)
        width = embedding_dim * 5 + 1
        self.cross_network = CrossNetwork(width, cross_layers)
        self.deep_network = DeepNetwork(width, deep_hidden_units)
        self.output_layer = nn.Linear(width + deep_hidden_units[-1], 1)
        for emb in (self.user_embedding, self.item_embedding, self.gender_embedding, self.occupation_embedding,
                    self.movie_embedding):
            xavier_normal_(emb.weight.data)

    def forward(self, x):
        z = K.stacked_features(x, self.user_embedding.weight, self.item_embedding.weight, self.gender_embedding.weight,
                               self.occupation_embedding.weight, self.movie_embedding.weight)
        both = torch.cat((self.cross_network(z), self.deep_network(z)), dim=1)
        return torch.sigmoid(self.output_layer(both))

    def recommendation(self, num_users, user_item, k):
        return K.topk_per_user(self, num_users, user_item, k)
