"""Device-side input preparation restated with numpy.  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

* ``philox4x32_10``: the published Philox4x32-10 counter RNG (Salmon et al., "Parallel random numbers: as easy as
  1, 2, 3", SC'11; Random123), pinned by its known-answer vectors in tests/test_oracle_golden.py.
* ``negative_draws_philox``: rs_sample_negatives' stream.  The ACCEPTANCE RULE is the reference's
  (sampler/sampler.py:21-27: redraw while the pair is excluded); the random stream is NOT python `random` (a
  sequential generator cannot be replayed by a parallel sampler), so parity with the reference is on the rule and on the
  distribution, not on the values: "parity unpinned" for the values, by construction.
* ``assemble_features``: data/reader.py:98-101, pinned against the reference's own ``MovieLens100K.feature`` in
  tests/golden/features.npz.
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(counter)[..., i].astype(np.uint64) for i in range(4)]
    k0 = np.asarray(key)[..., 0].astype(np.uint64)
    k1 = np.asarray(key)[..., 1].astype(np.uint64)
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _MASK, p1 >> np.uint64(32), p1 & _MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(_W0)) & _MASK
        k1 = (k1 + np.uint64(_W1)) & _MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def negative_draws_philox(num_user, num_item, excluded, num_negatives, seed, epoch=0, max_blocks=1 << 14):
    """(users, items) int64 arrays, user-major; `excluded` is a set of (user, item) pairs."""
    total = num_user * num_negatives
    users = np.repeat(np.arange(num_user, dtype=np.int64), num_negatives)
    items = np.full(total, -1, dtype=np.int64)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    s = np.arange(total, dtype=np.uint64)
    pending = np.arange(total)
    for blk in range(max_blocks):
        if pending.size == 0:
            break
        ctr = np.stack([(s[pending] & _MASK), (s[pending] >> np.uint64(32)), np.full(pending.size, blk, np.uint64),
                        np.full(pending.size, epoch, np.uint64)], axis=-1).astype(np.uint32)
        r = philox4x32_10(ctr, key)
        cand = (r.astype(np.uint64) * np.uint64(num_item)) >> np.uint64(32)
        for n, slot in enumerate(pending):
            for j in range(4):
                if (int(users[slot]), int(cand[n, j])) not in excluded:
                    items[slot] = int(cand[n, j])
                    break
        pending = pending[items[pending] < 0]
    return users, items


def assemble_features(users, items, user_feat, item_feat):
    """[user, item, user_feat[user], item_feat[item]] as float32 (B, 2 + FU + FI)."""
    users, items = np.asarray(users), np.asarray(items)
    return np.concatenate([users[:, None].astype(np.float32), items[:, None].astype(np.float32),
                           np.asarray(user_feat, np.float32)[users], np.asarray(item_feat, np.float32)[items]], axis=1)
