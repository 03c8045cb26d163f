"""Catalogue ranking and the ranking / classification metrics, restated with numpy and plain loops.
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Pinned by tests/test_oracle_golden.py against tests/golden/next_*.npz, recorded from the unmodified reference
(model/*.py recommendation(), evaluator/ranking.py, evaluator/evaluator.py) by tests/golden/make_golden_next.py.
"""
import math

import numpy as np


def rank_desc(scores, k):
    """positions of the k largest scores, descending; equal scores keep position order; NaN ranks first.
    This is what `torch.topk(scores, k, dim=0).indices` (model/deepfm.py:91-92) returns whenever scores are distinct."""
    s = np.asarray(scores, dtype=np.float32).reshape(-1)
    if k > s.size:
        raise RuntimeError("selected index k out of range")
    key = np.where(np.isnan(s), np.float32(np.inf), s)
    nan_first = np.isnan(s)
    order = np.lexsort((np.arange(s.size), -key.astype(np.float64), ~nan_first))
    return order[:k].astype(np.int64)


def rank_segments(scores, seg_start, k):
    seg_start = np.asarray(seg_start)
    return np.stack([rank_desc(scores[seg_start[s]:seg_start[s + 1]], k) for s in range(len(seg_start) - 1)])


def mf_scores(user_rows, item_rows):
    """model/mf.py:31: scores = U @ V^T (float64 here; the fp32 kernels must agree to 1e-5 relative)."""
    return np.asarray(user_rows, dtype=np.float64) @ np.asarray(item_rows, dtype=np.float64).T


# ---- evaluator/ranking.py:4-137, one definition per metric, scalar loops
def precision_recall_f1(actual, predicted, k):
    same = rec = real = 0
    for a, p in zip(actual, predicted):
        sa, sp = set(np.asarray(a).tolist()), set(np.asarray(p).tolist()[:k])
        same, rec, real = same + len(sa & sp), rec + len(sp), real + len(sa)
    pr, rc = same / rec, same / real
    return pr, rc, 2 * pr * rc / (pr + rc)


def apk(a_raw, p, k):
    """AP divides by len(actual) counted WITH duplicates (evaluator/ranking.py:58)."""
    aset = set(np.asarray(a_raw).tolist())
    hits, score = 0.0, 0.0
    for i, item in enumerate(np.asarray(p).tolist()[:k]):
        if item in aset:
            hits += 1.0
            score += hits / (i + 1.0)
    return score / len(a_raw)


def ndcg(a, p, k):
    aset = set(np.asarray(a).tolist())
    rel = [1 if item in aset else 0 for item in np.asarray(p).tolist()]

    def dcg(r):
        return sum((2 ** v - 1) / math.log2(i + 2) for i, v in enumerate(r[:k]))
    ideal = dcg(sorted(rel, reverse=True))
    return dcg(rel) / ideal if ideal > 0 else 0


def rr(a, p):
    aset = set(np.asarray(a).tolist())
    for i, item in enumerate(np.asarray(p).tolist()):
        if item in aset:
            return 1.0 / (i + 1)
    return 0.0


def ranking_metrics(actual, predicted, k):
    """[precision, recall, f1, MAP, mean NDCG, MRR] in the order ranking_eval prints them (ranking.py:127-137)."""
    pr, rc, f1 = precision_recall_f1(actual, predicted, k)
    return [pr, rc, f1, float(np.mean([apk(a, p, k) for a, p in zip(actual, predicted)])),
            float(np.mean([ndcg(a, p, k) for a, p in zip(actual, predicted)])),
            float(np.mean([rr(a, p) for a, p in zip(actual, predicted)]))]


def binary_metrics(y_true, y_pred):
    """evaluator/evaluator.py:13-20: threshold at >= 0.5, then accuracy / precision / recall / F1 / ROC-AUC of the
    hard predictions (sklearn definitions; AUC of a single operating point = (1 + TPR - FPR) / 2)."""
    t = np.asarray(y_true).reshape(-1) > 0.5
    p = np.asarray(y_pred).reshape(-1) >= 0.5
    tp, tn = int((t & p).sum()), int((~t & ~p).sum())
    fp, fn = int((~t & p).sum()), int((t & ~p).sum())
    prec = tp / (tp + fp) if tp + fp else 0.0
    rec = tp / (tp + fn) if tp + fn else 0.0
    f1 = 2 * tp / (2 * tp + fp + fn) if 2 * tp + fp + fn else 0.0
    fpr = fp / (fp + tn) if fp + tn else 0.0
    return [(tp + tn) / t.size, prec, rec, f1, 0.5 * (1 + rec - fpr)]
