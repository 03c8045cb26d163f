"""Field-aware FM over the six MovieLens features and two fields (user / item) -- drop-in for reference
model/ffm.py:7-98, including its quirk: the scalar cross term is added to every raw feature BEFORE the linear
layer, i.e. logit = u1 + i1 + W.(x[:,2:] + cross) + b  (model/ffm.py:84-86)."""
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from . import _blocks as K

_FEATS = ("age", "gender", "occupation", "movie", "userid", "itemid")
_FIELD_OF = (0, 0, 0, 1, 0, 1)                      # age, gender, occupation, userid -> user field; movie, itemid -> item
_ROWS = {"age": 1, "gender": 2, "occupation": 21, "movie": 19, "userid": 943, "itemid": 1682}
_SLOT = {"age": (K.AGE[0], K.AGE[1], K.KIND_BAG), "gender": (K.GENDER[0], K.GENDER[1], K.KIND_BAG),
         "occupation": (K.OCC[0], K.OCC[1], K.KIND_BAG), "movie": (K.GENRE[0], K.GENRE[1], K.KIND_BAG),
         "userid": (K.COL_USER, 1, K.KIND_ID), "itemid": (K.COL_ITEM, 1, K.KIND_ID)}


class FFM(nn.Module):
    def __init__(self, num_feature: int, num_vector: int):
        super().__init__()
        for f in _FEATS:                                            # registration order = reference state_dict order
            for fld in ("user", "item"):
                setattr(self, f"{f}_{fld}", nn.Embedding(_ROWS[f], num_vector))
        self.user = nn.Embedding(943, 1)
        self.item = nn.Embedding(1682, 1)
        self.linear = nn.Linear(num_feature, 1, True)
        for f in _FEATS:
            for fld in ("user", "item"):
                xavier_normal_(getattr(self, f"{f}_{fld}").weight.data)
        xavier_normal_(self.user.weight.data)
        xavier_normal_(self.item.weight.data)

    def forward(self, feature_vector: torch.Tensor) -> torch.Tensor:
        x = feature_vector
        slots, tables = [], []
        for f in _FEATS:                                            # T[b, feature, field, :]
            for fld in ("user", "item"):
                slots.append(_SLOT[f])
                tables.append(getattr(self, f"{f}_{fld}").weight)
        T = K.XEmbed.apply(x, tuple(slots), *tables)                # (B, 12, D)
        cross = K.FFMDense.apply(T.view(T.shape[0], 6, 2, T.shape[2]), _FIELD_OF)
        uid, iid = K.ops.xcol_to_ids(x, K.COL_USER), K.ops.xcol_to_ids(x, K.COL_ITEM)
        lin = self.linear(x[:, 2:] + cross.unsqueeze(1))
        return torch.sigmoid(K.lookup(self.user.weight, uid) + K.lookup(self.item.weight, iid) + lin)

    def recommendation(self, num_users, user_item, k):
        return K.topk_per_user(self, num_users, user_item, k)
