"""CUDA-graph capture of a whole train step (forward -> loss -> backward -> fused row update -> dense optimizer).

The step of the small-row configs is launch bound (C5 MF: ~0.3 ms of kernels inside a ~1.1 ms step), so the launch
sequence -- our C-ABI kernels and torch's elementwise/optimizer kernels alike -- is recorded once and replayed.
Everything in the step must be shape-static and free of host synchronisation, which holds for the single-GPU fused
path (`FusedRowOptimizer` + `nfield` models, drop-in modules with torch optimizers).  The row-sharded multi-GPU path
syncs for its all-to-all split sizes and is not captured.
"""
import torch


class GraphedTrainStep:
    """step(*inputs, rating) -> (predictions, loss) with identical semantics to Trainer.train_loop.

    If the model has already been trained eagerly, drop every tensor that still references that autograd graph (e.g.
    `trainer.train_loss`, `trainer.predictions_train`) first: autograd binds a parameter's AccumulateGrad node to the
    stream it was created on, and a node from the legacy default stream cannot be reused during capture.

    The first `warmup` calls run eagerly (they are real train steps on their own inputs and also let every library
    and cache initialise); the next call captures the step on static input buffers and replays it.  From then on
    inputs are copied into the static buffers, the graph is replayed, and the static outputs are returned (valid
    until the next call).
    """

    def __init__(self, model, loss_fn, optimizer, warmup=2, also=()):
        """also: further (model, loss_fn, optimizer) triples stepped on the SAME inputs inside the same graph (the C2
        benchmark's FFM + FM pair, which share one dedup / one exchange plan per batch); the first triple's
        (predictions, loss) are returned."""
        self.model, self.loss_fn, self.optimizer, self.warmup = model, loss_fn, optimizer, warmup
        self.jobs = [(model, loss_fn, optimizer)] + [tuple(j) for j in also]
        self.graph, self.calls, self.side = None, 0, None
        self.static_in = self.static_rating = self.pred = self.loss = None
        self.all_losses = None

    def _eager(self, inputs, rating):
        outs = []
        for model, loss_fn, optimizer in self.jobs:
            from .trainer.trainer import fused_bce_step
            model.train()
            optimizer.zero_grad()
            fused = fused_bce_step(model, loss_fn, inputs, rating)
            if fused is not None:
                pred, loss = fused
            else:
                pred = model(*inputs)
                loss = loss_fn(pred, rating)
                loss.backward()
            optimizer.step()
            outs.append((pred, loss))
        self.all_losses = [l for _, l in outs]
        return outs[0]

    def _check_capturable(self):
        """Adam's bias correction is computed on the host from a python step counter and handed to the kernels by value
        (rs_segment_update / rs_adam_dense), so a captured Adam step would replay with the correction frozen at the
        capture step.  Refuse instead of training silently wrong; SGD (and torch optimizers built with
        capturable=True, whose step lives on the device) capture fine."""
        from .optim import DenseAdam, FusedRowOptimizer
        opts = []
        for _, _, opt in self.jobs:
            if isinstance(opt, FusedRowOptimizer):
                if opt.kind == "adam":
                    raise RuntimeError("GraphedTrainStep: FusedRowOptimizer(kind='adam') keeps its step count on the host and "
                                       "cannot be captured; run it eagerly or use kind='sgd'")
                opts.append(opt.dense)
            else:
                opts.append(opt)
        for o in opts:
            if isinstance(o, DenseAdam):
                raise RuntimeError("GraphedTrainStep: DenseAdam keeps its step count on the host and cannot be captured")
            if isinstance(o, (torch.optim.Adam, torch.optim.AdamW)) and not all(g.get("capturable") for g in o.param_groups):
                raise RuntimeError("GraphedTrainStep: torch Adam must be built with capturable=True to be captured")

    def _capture(self, inputs, rating):
        self._check_capturable()
        self.static_in = [t.clone() for t in inputs]
        self.static_rating = rating.clone()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            pred, loss = self._eager(self.static_in, self.static_rating)
        self.pred, self.loss = pred, loss

    def __call__(self, *inputs, rating):
        self.calls += 1
        if self.graph is None and self.calls <= self.warmup:
            # eager steps run on a side stream: autograd remembers the stream a parameter's AccumulateGrad node was
            # created on, and a node created on the legacy default stream cannot be used while capturing
            if self.side is None:
                self.side = torch.cuda.Stream()
            cur = torch.cuda.current_stream()
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                out = self._eager(inputs, rating)
            cur.wait_stream(self.side)
            return out
        if self.graph is None:
            self._capture(inputs, rating)
            # the capture itself does not execute the step; fall through to one replay on these inputs
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.static_rating.copy_(rating, non_blocking=True)
        self.graph.replay()
        return self.pred, self.loss
