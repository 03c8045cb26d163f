"""OPTIONAL device-side negative sampling and feature assembly.

``DeviceSampler`` has the reference Sampler's surface (sampler/sampler.py:11-27: ``negative_sampling(num_user, num_item,
excluded_pairs, num_negatives, device)`` -> (users, items, zeros); negatives accumulate on the instance across calls)
and the same acceptance rule, but draws from a counter-based Philox stream on the GPU instead of python's global
``random``.  The values therefore differ from the reference's for any seed -- scripts that must replay the reference
bit for bit keep ``sampler.Sampler``; this class is for sampling fresh negatives every epoch without a host loop.
``features`` builds the (B, 45) model input from id tensors and the two side-feature tables (data/reader.py:98-101).
"""
import numpy as np
import torch

from .. import ops


class DeviceSampler:
    def __init__(self, seed=0):
        self.seed, self.epoch = int(seed), 0
        self.negative_users, self.negative_items = [], []      # lists of CUDA tensors, one per call
        self._keys = None

    def _excluded_keys(self, excluded_pairs, num_item, device):
        """ascending int64 keys user * num_item + item on the device; cached for the same set object."""
        if self._keys is not None and self._keys[0] is excluded_pairs and self._keys[1] == (num_item, str(device)):
            return self._keys[2]
        if isinstance(excluded_pairs, torch.Tensor):
            pairs = excluded_pairs.to(device=device, dtype=torch.int64).view(-1, 2)
            keys = torch.sort(pairs[:, 0] * num_item + pairs[:, 1]).values
        else:
            arr = np.fromiter((int(u) * num_item + int(i) for u, i in excluded_pairs), dtype=np.int64, count=len(excluded_pairs))
            arr.sort()
            keys = torch.from_numpy(arr).to(device)
        self._keys = (excluded_pairs, (num_item, str(device)), keys)
        return keys

    def negative_sampling(self, num_user, num_item, excluded_pairs, num_negatives, device="cuda"):
        keys = self._excluded_keys(excluded_pairs, num_item, torch.device(device))
        users, items = ops.sample_negatives(keys, num_user, num_item, num_negatives, self.seed, self.epoch)
        self.epoch += 1
        self.negative_users.append(users)
        self.negative_items.append(items)
        users, items = torch.cat(self.negative_users), torch.cat(self.negative_items)
        return users, items, torch.zeros(users.numel(), device=users.device)

    @staticmethod
    def features(users, items, user_feat, item_feat):
        """[user, item, user_feat[user], item_feat[item]] fp32 -- what ``data.feature(...)`` + ``.values`` yields."""
        return ops.assemble_features(users, items, user_feat, item_feat)
