"""Top-k ranking metrics -- mirror of reference evaluator/ranking.py:4-137.

Same constructor and methods (``precision_recall_f1``, ``apk``, ``mapk``, ``dcg``, ``ndcg``, ``mean_ndcg``, ``rr``,
``mrr``, ``ranking_eval``) and the same definitions, quirks included: precision/recall de-duplicate both lists as sets
(ranking.py:24-31); AP counts every hit position of the first k predictions and divides by ``len(actual)``
(ranking.py:52-58); NDCG takes the relevance of the *whole* predicted list and truncates only inside ``dcg``
(ranking.py:85-93); RR scans the whole list (ranking.py:108-111).  The reference tests membership with python
``in`` per element; here each user is one ``np.isin`` and the running sums keep the reference's left-to-right order
(``np.cumsum``), so the values are the same doubles.  This is host bookkeeping after training, not the GPU hot path.
"""
import numpy as np


def _hits(actual, predicted):
    predicted = np.asarray(predicted)
    if predicted.size == 0:
        return np.zeros(0, dtype=bool)
    return np.isin(predicted, np.asarray(actual))


class Ranking:
    def __init__(self, real_list, rec_list, k):
        self.actual = real_list
        self.predicted = rec_list
        self.k = k

    def precision_recall_f1(self):
        same = rec = real = 0
        for a, p in zip(self.actual, self.predicted):
            relevant, recommended = np.unique(np.asarray(a)), np.unique(np.asarray(p)[:self.k])
            same += int(np.isin(recommended, relevant).sum())
            rec += recommended.size
            real += relevant.size
        precision = same / (rec * 1.0)
        recall = same / (real * 1.0)
        return precision, recall, 2 * (precision * recall) / (precision + recall)

    @staticmethod
    def apk(actual, predicted, k):
        hit = _hits(actual, np.asarray(predicted)[:k])
        if not hit.any():
            return 0.0 / len(actual)
        pos = np.flatnonzero(hit)
        terms = np.arange(1, pos.size + 1, dtype=np.float64) / (pos + 1.0)
        return float(np.cumsum(terms)[-1]) / len(actual)

    def mapk(self):
        return np.mean([self.apk(a, p, self.k) for a, p in zip(self.actual, self.predicted)])

    @staticmethod
    def dcg(relevance_scores, k):
        rel = np.asarray(relevance_scores)[:k]
        return np.sum((2 ** rel - 1) / np.log2(np.arange(1, len(rel) + 1) + 1))

    def ndcg(self, actual, predicted, k):
        rel = _hits(actual, predicted).astype(np.int64)
        dcg_score = self.dcg(rel, k)
        idcg_score = self.dcg(np.sort(rel)[::-1], k)
        return dcg_score / idcg_score if idcg_score > 0 else 0

    def mean_ndcg(self):
        return np.mean([self.ndcg(a, p, self.k) for a, p in zip(self.actual, self.predicted)])

    @staticmethod
    def rr(actual, predicted):
        pos = np.flatnonzero(_hits(actual, predicted))
        return 1.0 / (int(pos[0]) + 1) if pos.size else 0.0

    def mrr(self):
        return np.mean([self.rr(a, p) for a, p in zip(self.actual, self.predicted)])

    def ranking_eval(self):
        precision, recall, f1 = self.precision_recall_f1()
        k = self.k
        print(f"\n    - Precision@{k}:  {precision}\n    - Recall@{k}:  {recall}\n    - F1 Score@{k}:  {f1}\n"
              f"    - MAP@{k}: {self.mapk()}\n    - Mean NDCG@{k}: {self.mean_ndcg()}\n    - MRR: {self.mrr()}\n")
