"""Classification metrics -- mirror of reference evaluator/evaluator.py:8-20.

``Evaluator.eval(y_true, y_pred)`` returns ``[accuracy, precision, recall, f1, roc_auc]`` exactly as the reference's
sklearn calls do for its inputs (labels in {0, 1}; probabilities thresholded at ``>= 0.5`` *before* the AUC,
evaluator/evaluator.py:17-19), but from four confusion counts reduced on the device: one small host read per call
instead of copying both vectors to the host three times per epoch (trainer/trainer.py:116-120).
"""
import torch


def binary_metrics(y_true, y_pred):
    """[accuracy, precision, recall, f1, auc] with sklearn's conventions for hard 0/1 predictions."""
    t = (y_true.detach().reshape(-1) > 0.5)
    p = (y_pred.detach().reshape(-1) >= 0.5)
    tp = (t & p).sum().double()
    tn = (~t & ~p).sum().double()
    fp = (~t & p).sum().double()
    fn = (t & ~p).sum().double()
    n = tp + tn + fp + fn
    zero = torch.zeros((), dtype=torch.float64, device=tp.device)
    acc = (tp + tn) / n
    prec = torch.where(tp + fp > 0, tp / (tp + fp).clamp(min=1), zero)       # sklearn zero_division -> 0
    rec = torch.where(tp + fn > 0, tp / (tp + fn).clamp(min=1), zero)
    f1 = torch.where(2 * tp + fp + fn > 0, 2 * tp / (2 * tp + fp + fn).clamp(min=1), zero)
    fpr = torch.where(fp + tn > 0, fp / (fp + tn).clamp(min=1), zero)
    auc = 0.5 * (1.0 + rec - fpr)          # the ROC of a hard classifier has a single interior point
    return [float(v) for v in torch.stack([acc, prec, rec, f1, auc]).cpu()]


class Evaluator:
    def __init__(self):
        self.array = []

    @staticmethod
    def eval(y_true, y_pred):
        return binary_metrics(y_true, y_pred)
