"""Host-side negative sampling -- mirror of reference sampler/sampler.py:11-48.

Bit-exact obligation: under the same ``random.seed`` the (user, item) stream must equal the
reference's, so this stays on the host and consumes python's global ``random`` in exactly the same order: for each
user in turn, ``num_negatives`` accepted draws of ``random.randint(0, num_item - 1)``, redrawing while the pair is in
``excluded_pairs``.  Like the reference, negatives ACCUMULATE on the instance across calls.
"""
import random

import torch


class Sampler:
    def __init__(self):
        self.negative_users = []
        self.negative_items = []

    def _draw(self, num_user, num_item, excluded_pairs, num_negatives):
        top = num_item - 1
        for user in range(num_user):
            accepted = 0
            while accepted < num_negatives:
                item = random.randint(0, top)
                if (user, item) in excluded_pairs:
                    continue
                self.negative_users.append(user)
                self.negative_items.append(item)
                accepted += 1

    def negative_sampling(self, num_user, num_item, excluded_pairs, num_negatives, device="cpu"):
        """-> (users int64, items int64, zeros float32), each of length len(all negatives drawn so far)."""
        self._draw(num_user, num_item, excluded_pairs, num_negatives)
        users = torch.tensor(self.negative_users).to(device)
        items = torch.tensor(self.negative_items).to(device)
        return users, items, torch.zeros(len(self.negative_users)).to(device)

    def negative_sampling2(self, num_user, num_item, excluded_pairs, num_negatives):
        """Same stream as a DataFrame with columns user_id, item_id, rating (= 0)."""
        import pandas as pd
        self._draw(num_user, num_item, excluded_pairs, num_negatives)
        return pd.DataFrame({"user_id": self.negative_users, "item_id": self.negative_items,
                             "rating": [0] * len(self.negative_users)})
