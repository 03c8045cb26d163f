"""NeuralCF (GMF * MLP tower) -- drop-in for reference model/neuralcf.py:7-73."""
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from . import _blocks as K


class NeuralCF(nn.Module, K.FusedRows):
    def __init__(self, num_user, num_item, mf_dim, layers):
        super().__init__()
        self.GMF_Embedding_User = nn.Embedding(num_user, mf_dim)
        self.GMF_Embedding_Item = nn.Embedding(num_item, mf_dim)
        self.MLP_Embedding_User = nn.Embedding(num_user, int(layers[0] / 2))
        self.MLP_Embedding_Item = nn.Embedding(num_item, int(layers[0] / 2))
        for emb in (self.GMF_Embedding_User, self.GMF_Embedding_Item, self.MLP_Embedding_User, self.MLP_Embedding_Item):
            xavier_normal_(emb.weight.data)
        self.dnn_network = nn.ModuleList([nn.Linear(a, b) for a, b in zip(layers[:-1], layers[1:])])
        self.relu = nn.ReLU()
        self.linear = nn.Linear(layers[-1], mf_dim)
        self.linear2 = nn.Linear(2 * mf_dim, 1)
        self.sigmoid = nn.Sigmoid()

    def _fused_groups(self):
        return [[self.GMF_Embedding_User, self.GMF_Embedding_Item], [self.MLP_Embedding_User, self.MLP_Embedding_Item]]

    def forward(self, user_indices, item_indices):
        gmf = K.pair_lookup(self.GMF_Embedding_User.weight, self.GMF_Embedding_Item.weight, user_indices, item_indices, "had2")
        x = K.pair_lookup(self.MLP_Embedding_User.weight, self.MLP_Embedding_Item.weight, user_indices, item_indices, "concat")
        for layer in self.dnn_network:
            x = self.relu(layer(x))
        vector = torch.cat([gmf, self.linear(x)], dim=1)
        return self.sigmoid(self.linear2(vector))

    def recommendation(self, num_users, num_items):
        device = next(self.parameters()).device
        items = torch.arange(num_items, device=device)

        def score(u0, u1):
            users = torch.arange(u0, u1, device=device).repeat_interleave(num_items)
            return self.forward(users, items.repeat(u1 - u0))
        return K.rank_catalogue(score, num_users, num_items, num_items)
