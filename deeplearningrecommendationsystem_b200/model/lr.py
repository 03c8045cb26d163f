"""Logistic regression over [uid, iid, 43 side features] -- drop-in for reference model/lr.py:11-37."""
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from . import _blocks as K


class LogisticRegression(nn.Module):
    def __init__(self, num_users, num_items, num_feature: int):
        super().__init__()
        self.user = nn.Embedding(num_users, 1)
        self.item = nn.Embedding(num_items, 1)
        self.linear = nn.Linear(num_feature, 1, True)
        xavier_normal_(self.user.weight.data)
        xavier_normal_(self.item.weight.data)

    def forward(self, feature_vector: torch.Tensor) -> torch.Tensor:
        return torch.sigmoid(K.first_order(self.user, self.item, self.linear, feature_vector))

    def recommendation(self, num_users, user_item, k):
        return K.topk_per_user(self, num_users, user_item, k)
