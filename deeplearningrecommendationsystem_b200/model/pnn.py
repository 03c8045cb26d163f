"""Product-based NN -- drop-in for reference model/pnn.py:8-143 (DNN, ProductLayers, PNN).

Quirks kept: z is unsqueezed to (1, B, 6D) so the tower runs 3-D and the output is .view(-1, 1); "out" mode
collapses the batch (S^T S with S = sum_f e_f) and therefore only broadcasts when B == embed_dim -- any other
batch raises the same RuntimeError as the reference.  User/item tables are hard-coded to
943 / 1682 rows (model/pnn.py:87-88)."""
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from . import _blocks as K


class DNN(nn.Module):
    def __init__(self, hidden_units):
        super().__init__()
        self.dnn_network = nn.ModuleList([nn.Linear(a, b) for a, b in zip(hidden_units[:-1], hidden_units[1:])])
        self.relu = nn.ReLU()

    def forward(self, x):
        for layer in self.dnn_network:
            x = self.relu(layer(x))
        return x


class ProductLayers(nn.Module):
    def __init__(self, num_feature, embed_dim, hidden_units, model="in"):
        super().__init__()
        self.model = model
        self.linear1 = nn.Linear(num_feature * embed_dim, hidden_units[0])
        if model == "in":
            self.linear2 = nn.Linear(int(num_feature * (num_feature - 1) / 2), hidden_units[0])
        elif model == "out":
            self.linear2 = nn.Linear(embed_dim, hidden_units[0])

    def forward(self, feature_embed):
        """feature_embed: (B, F, D) tensor, or the reference's list of F (B, D) tensors."""
        E = torch.stack(list(feature_embed), dim=1) if isinstance(feature_embed, (list, tuple)) else feature_embed
        z = E.flatten(1).unsqueeze(0)
        if self.model == "in":
            p = K.inner_products(E)
        elif self.model == "out":
            s = E.sum(dim=1)
            p = K.OuterPooled.apply(s) if s.is_cuda and s.shape[1] <= 128 else torch.matmul(s.T, s)
        return self.linear1(z) + self.linear2(p)


class PNN(nn.Module):
    def __init__(self, embed_dim, hidden_units, model="in"):
        super().__init__()
        self.user_embed = nn.Embedding(943, embed_dim)
        self.item_embed = nn.Embedding(1682, embed_dim)
        self.age_embed = nn.Embedding(1, embed_dim)
        self.gender_embed = nn.Embedding(2, embed_dim)
        self.occupation_embed = nn.Embedding(21, embed_dim)
        self.movie_embed = nn.Embedding(19, embed_dim)
        for emb in (self.user_embed, self.item_embed, self.age_embed, self.gender_embed, self.occupation_embed, self.movie_embed):
            xavier_normal_(emb.weight.data)
        self.product = ProductLayers(6, embed_dim, hidden_units, model)
        self.dnn = DNN(hidden_units)
        self.output = nn.Linear(hidden_units[-1], 1)

    def forward(self, x):
        E = K.XEmbed.apply(x, K.six_slots(), self.user_embed.weight, self.item_embed.weight, self.age_embed.weight,
                           self.gender_embed.weight, self.occupation_embed.weight, self.movie_embed.weight)
        h = self.dnn(self.product(E))
        return torch.sigmoid(self.output(h)).view(-1, 1)

    def recommendation(self, num_users, user_item, k):
        # "out" pools the outer product over the batch (model/pnn.py:72), so a user's scores depend on exactly which
        # rows share the forward: keep one forward per user there
        return K.topk_per_user(self, num_users, user_item, k, batched=self.product.model == "in")
