"""Train / validate / test step protocol -- mirror of reference trainer/trainer.py:8-146.

Same public surface: ``Trainer(model, loss_fn, optimizer)``; ``train_loop(*args, train_rating=)``,
``valid_loop``, ``test_loop`` dispatch on the number of positional inputs (1 -> ``model(x)``, 2 -> ``model(a, b)``,
anything else -> ValueError); the ``*_loop2(matrix, mask)`` AutoRec variants; results are left on the instance in
``predictions_*``, ``*_loss`` and ``*_rating``.  ``optimizer`` only needs ``zero_grad()`` and ``step()``, so a
``FusedRowOptimizer`` drops in.  ``model_eval`` computes the reference Evaluator's five metrics
(evaluator/evaluator.py:13-20: predictions thresholded at >= 0.5 *before* AUC) on the device, one host read per split.
"""
import os

import torch

from ..evaluator.evaluator import binary_metrics

# Out-of-range ids never fault on the device (they are clamped and a status bit is set); the reference raises
# IndexError from nn.Embedding.  The Trainer turns the bit back into that IndexError: on every valid/test pass and
# metrics() call (per-epoch, they read results back anyway) and every RS_CHECK_EVERY train steps (one 4-byte read).
CHECK_EVERY = int(os.environ.get("RS_CHECK_EVERY", "64"))


def _check_ids(model):
    try:
        dev = next(model.parameters()).device
    except StopIteration:
        return
    if dev.type != "cuda" or torch.cuda.is_current_stream_capturing():
        return
    from .. import ops
    ops.check_status(dev)


_binary_metrics = binary_metrics


def fused_bce_step(model, loss_fn, args, rating):
    """forward -> sigmoid -> BCELoss(mean) -> backward with the sigmoid, the loss, its mean and both of their backward
    passes in ONE kernel pair (rs_sigmoid_bce: same op sequence as autograd, fixed-order mean), for models that expose
    their pre-sigmoid logit through ``train_logit`` -> (cross, bias parameter | None, output shape) with logit = cross + bias
    (the nfield FM / FFM / MF models) under a plain ``nn.BCELoss()``.
    Returns (predictions, loss) exactly as ``loss_fn(model(*args), rating)`` would (loss detached), or None when the
    combination does not apply -- the caller then runs the generic path.  RS_FUSED_BCE=0 disables it."""
    if os.environ.get("RS_FUSED_BCE", "1") != "1" or type(loss_fn) is not torch.nn.BCELoss:
        return None
    if loss_fn.reduction != "mean" or loss_fn.weight is not None:
        return None
    fn = getattr(model, "train_logit", None)
    if fn is None or not torch.is_grad_enabled() or not args[0].is_cuda or rating.dtype != torch.float32:
        return None
    if rating.numel() != args[0].shape[0] or torch.is_tensor(rating) and rating.requires_grad:
        return None
    from .. import ops
    cross, bias, shape = fn(*args)                    # logit = cross + bias (bias may be None)
    if tuple(rating.shape) != tuple(shape):           # BCELoss would raise / warn: let it
        raise ValueError(f"Using a target size ({tuple(rating.shape)}) that is different to the input size ({tuple(shape)}) is deprecated. "
                         "Please ensure they have the same size.")
    want_gsum = bias is not None and bias.requires_grad
    out = ops.sigmoid_bce(cross.detach(), rating, bias=None if bias is None else bias.detach(), want_gsum=want_gsum)
    pred, loss, g = out[:3]
    cross.backward(g)
    if want_gsum:                                     # d loss / d bias = sum of d loss / d logit, reduced inside the loss kernel
        gb = out[3].view_as(bias)
        bias.grad = gb if bias.grad is None else bias.grad + gb
    return pred.view(shape), loss


class Trainer:
    def __init__(self, model, loss_fn, optimizer):
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.train_loss = self.valid_loss = self.test_loss = None
        self._steps = 0
        self.predictions_train = self.predictions_valid = self.predictions_test = None
        self.train_rating = self.valid_rating = self.test_rating = None

    def _forward(self, args, who):
        if len(args) not in (1, 2):
            raise ValueError(f"Invalid number of arguments provided to {who}")
        return self.model(*args)

    def train_loop(self, *args, train_rating):
        self.model.train()
        self.optimizer.zero_grad()
        if len(args) not in (1, 2):
            raise ValueError("Invalid number of arguments provided to train_loop")
        fused = fused_bce_step(self.model, self.loss_fn, args, train_rating)
        if fused is not None:
            self.predictions_train, self.train_loss = fused
        else:
            self.predictions_train = self._forward(args, "train_loop")
            self.train_loss = self.loss_fn(self.predictions_train, train_rating)
            self.train_loss.backward()
        self.optimizer.step()
        self.train_rating = train_rating
        self._steps += 1
        if CHECK_EVERY > 0 and self._steps % CHECK_EVERY == 0:
            _check_ids(self.model)

    def _eval(self, args, rating, who):
        self.model.eval()
        with torch.no_grad():
            pred = self._forward(args, who)
            loss = self.loss_fn(pred, rating)
        _check_ids(self.model)
        return pred, loss

    def valid_loop(self, *args, valid_rating):
        self.predictions_valid, self.valid_loss = self._eval(args, valid_rating, "valid_loop")
        self.valid_rating = valid_rating

    def test_loop(self, *args, test_rating):
        self.predictions_test, self.test_loss = self._eval(args, test_rating, "test_loop")
        self.test_rating = test_rating

    # masked single-input variants (AutoRec scripts)
    def train_loop2(self, train_matrix, mask):
        self.model.train()
        self.optimizer.zero_grad()
        self.predictions_train = self.model(train_matrix)[mask]
        self.train_rating = train_matrix[mask]
        self.train_loss = self.loss_fn(self.predictions_train, self.train_rating)
        self.train_loss.backward()
        self.optimizer.step()

    def _eval2(self, matrix, mask):
        self.model.eval()
        with torch.no_grad():
            pred = self.model(matrix)[mask]
            rating = matrix[mask]
            return pred, rating, self.loss_fn(pred, rating)

    def valid_loop2(self, valid_matrix, mask):
        self.predictions_valid, self.valid_rating, self.valid_loss = self._eval2(valid_matrix, mask)

    def test_loop2(self, test_matrix, mask):
        self.predictions_test, self.test_rating, self.test_loss = self._eval2(test_matrix, mask)

    def metrics(self):
        """{'train'|'valid'|'test': [acc, precision, recall, f1, auc]} for the splits that have run."""
        out = {}
        _check_ids(self.model)
        for split in ("train", "valid", "test"):
            pred, rating = getattr(self, f"predictions_{split}"), getattr(self, f"{split}_rating")
            if pred is not None:
                out[split] = _binary_metrics(rating, pred)
        return out

    def model_eval(self, epoch):
        m = self.metrics()
        names = ["Accuracy", "Precision", "Recall", "F1 Score", "ROC AUC Score"]
        title = {"train": "Training", "valid": "Valid", "test": "Test"}
        lines = [f"Epoch {epoch + 1}:"]
        for split in m:
            lines.append(f"  - {title[split]} Loss: {getattr(self, split + '_loss').item()}")
        for k, name in enumerate(names):
            for split in m:
                lines.append(f"  - {title[split]} {name}: {m[split][k]}")
        print("\n".join(lines))
