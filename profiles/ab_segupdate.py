"""A/B timing of rs_segment_update on the C2 FFM shape (1.7 M lookups, 416-float rows): fused SGD into the table vs
RS_UPD_GRAD into a compact buffer (plain pointer / routed / routed through the pusher warp).  One GPU."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearningrecommendationsystem_b200 import ops  # noqa: E402

CRITEO = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992, 5461306, 10, 5652,
          2173, 4, 7046547, 18, 15, 286181, 105, 142572]
cards = [min(c, 1 << 17) for c in CRITEO] if "--light" in sys.argv else CRITEO
F, W, B = 26, 416, 65536
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
ids = torch.stack([torch.randint(0, c, (B,), generator=g, device=dev) for c in cards], dim=1).contiguous()
offs = [sum(cards[:i]) for i in range(F)]
total = sum(cards)
stash = torch.randn(B, F, W, device=dev)
scale = torch.randn(B, device=dev)
table = torch.zeros(total, W, device=dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


segs = ops.dedup_sort(ids, F, offs, total, max_width=W, reuse_workspace=False)
nu = segs.n_uniq
print("lookups", ids.numel(), "unique", nu)
print("SGD fused into table        %.3f ms" % timeit(lambda: ops.segment_update(segs, ops.RS_UPD_SGD, W, F, stash=stash, scale=scale, table=table, lr=0.01)))
blk = ops.dedup_sort(ids, F, offs, total, max_width=W, reuse_workspace=False)
blk = ops.block_segments(blk, nu, W)
grad = torch.zeros(ids.numel(), W, device=dev)
print("GRAD -> compact dense_grad  %.3f ms" % timeit(lambda: ops.segment_update(blk, ops.RS_UPD_GRAD, W, F, stash=stash, scale=scale, dense_grad=grad)))
routes = ops.make_routes([0, ids.numel()], [grad.data_ptr()], [0])
print("GRAD routed (direct stores) %.3f ms" % timeit(lambda: ops.segment_update(blk, ops.RS_UPD_GRAD, W, F, stash=stash, scale=scale, grad_routes=routes)))
os.environ["RS_PUSHER"] = "1"
print("GRAD routed (pusher warp)   %.3f ms" % timeit(lambda: ops.segment_update(blk, ops.RS_UPD_GRAD, W, F, stash=stash, scale=scale, grad_routes=routes)))
del os.environ["RS_PUSHER"]
print("SGD fused, relabelled segs  %.3f ms" % timeit(lambda: ops.segment_update(blk, ops.RS_UPD_SGD, W, F, stash=stash, scale=scale, table=grad, lr=0.01)))
