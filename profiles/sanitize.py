"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck) over the hand-written pipelines: the TMA +
mbarrier producer/consumer kernels (ffm_fwd_kernel, seg_stream_kernel<*>, shard_serve_tma_kernel), the tcgen05 kernels
(AFM / DIN / gemm_tn) and the peer-store path of the row-sharded step (virtual ranks on one GPU: the same kernels and
pointers as the multi-GPU step).  Sizes are tiny: the sanitizer slows kernels down by 10-100x.

    compute-sanitizer --tool memcheck  python profiles/sanitize.py
    compute-sanitizer --tool racecheck python profiles/sanitize.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearningrecommendationsystem_b200 import dist as rsdist, ops  # noqa: E402
from deeplearningrecommendationsystem_b200.nfield import FieldAFM, FieldFFM, FieldFM  # noqa: E402
from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer  # noqa: E402
from deeplearningrecommendationsystem_b200.trainer import Trainer  # noqa: E402

CRITEO = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992, 5461306, 10, 5652,
          2173, 4, 7046547, 18, 15, 286181, 105, 142572]
cards = [min(c, 500) for c in CRITEO]
g = torch.Generator().manual_seed(0)


def batch(B, seed=0):
    gg = torch.Generator().manual_seed(seed)
    ids = torch.stack([torch.randint(0, c, (B,), generator=gg) for c in cards], dim=1).cuda()
    return ids, (torch.rand(B, 1, generator=gg) < 0.3).float().cuda()


# 1. C2 step at its own shape (F = 26, D = 16): ffm_fwd_kernel + seg_stream_kernel<SGD> + chunk / combine kernels
for cls in (FieldFFM, FieldFM):
    m = cls(cards, 16, seed=1, device="cuda")
    tr = Trainer(m, torch.nn.BCELoss(), FusedRowOptimizer(m, torch.optim.SGD([m.bias], lr=0.1), lr=0.1))
    for k in range(2):
        tr.train_loop(*batch(600, k)[:1], train_rating=batch(600, k)[1])
print("c2 step ok", flush=True)


# 2. row-sharded step, 3 virtual ranks: shard_post / collect / serve (TMA and 128-bit), routed seg_stream<GRAD>, n_valid dedup
def rank_fn(fab):
    ex = rsdist.DeviceRowExchange(fab)
    ms = [cls(cards, 16, seed=2, device="cuda", sharded=True, exchange=ex) for cls in (FieldFFM, FieldFM)]
    trs = [Trainer(m, torch.nn.BCELoss(), FusedRowOptimizer(m, torch.optim.SGD([m.bias], lr=0.1), lr=0.1)) for m in ms]
    ids, y = batch(300, 10 + fab.rank)
    for _ in range(2):
        for tr in trs:
            tr.train_loop(ids, train_rating=y)
    return True


rsdist.ThreadFabric.run(3, rank_fn)
print("sharded step ok", flush=True)

# 3. tcgen05 kernels: AFM forward / backward (F = 39, D = 32, A = 64), DIN attention forward / backward, gemm_tn
B = 2 * 148 + 8
E = (torch.randn(B, 39, 32, generator=g) * 0.5).cuda()
W, b, h = (torch.randn(32, 64, generator=g) * 0.3).cuda(), torch.randn(64, generator=g).cuda(), torch.randn(64, 1, generator=g).cuda()
pooled, attw = ops.afm_fwd(E, W, b, h)
ops.afm_bwd(E, W, b, h, attw, torch.randn(B, 32, generator=g).cuda())
rows = (torch.randn(300, 101, 64, generator=g) * 0.5).cuda()
lin = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5).cuda()   # noqa: E731
vec = lambda o, i: ((torch.rand(o, generator=g) * 2 - 1) / i ** 0.5).cuda()      # noqa: E731
cw = [lin(128, 192), vec(128, 192), lin(64, 128), vec(64, 128), lin(1, 64), vec(1, 64)]
out, aw, stash = ops.din_fwd(rows, cw, True, impl="tc", want_stash=True)
ops.din_bwd_tc(rows, cw, True, torch.randn(300, 64, generator=g).cuda(), stash)
ops.gemm_tn(torch.randn(5000, 64, generator=g).cuda(), torch.randn(5000, 32, generator=g).cuda())
torch.cuda.synchronize()
ops.check_status()
print("tensor-core kernels ok", flush=True)
