"""The host-sync-free row-sharded step (dist.DeviceRowExchange + csrc/shard.cu) on ONE GPU.

The kernels only see pointers, so N ranks can be N python threads whose "peer" buffers are each other's tensors
(dist.ThreadFabric): the same kernels and launch order as the multi-GPU step, checked against the unsharded step on the
concatenated global batch.  (tests/test_dist_gpu.py runs the same comparison over real NVLink peers when >= 2 GPUs exist.)
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
CARDS = [3, 50, 7, 1000, 24, 12, 301, 5]


def _ops():
    from deeplearningrecommendationsystem_b200 import ops
    return ops


def _batch(rank, B, cards=CARDS):
    g = torch.Generator().manual_seed(500 + rank)
    ids = torch.stack([(torch.rand(B, generator=g) ** 2 * c).long().clamp_(max=c - 1) for c in cards], dim=1)
    return ids, (torch.rand(B, 1, generator=g) < 0.3).float()


@pytest.mark.parametrize("n,nv,rows", [(1000, 1000, 50), (1000, 377, 50), (4096, 1, 7), (513, 0, 9), (70000, 33333, 100000)])
def test_dedup_with_device_side_length(n, nv, rows):
    """rs_dedup_sort_ex(n_valid): a fixed-capacity list whose fill level is a device scalar gives exactly the segments
    of its valid prefix, and the segment-reduce over the full capacity touches only rows of that prefix."""
    ops = _ops()
    g = torch.Generator().manual_seed(n + nv)
    ids = torch.randint(0, rows, (n,), generator=g)
    ids[nv:] = 10 ** 12                                          # padding is garbage: must never be read as an id
    nvt = torch.tensor([nv], dtype=torch.int32).cuda()
    segs = ops.dedup_sort(ids.cuda(), 1, None, rows, max_width=8, n_valid=nvt)
    u, inv, cnt = torch.unique(ids[:nv], sorted=True, return_inverse=True, return_counts=True)
    assert segs.n_uniq == u.numel()
    assert torch.equal(segs.uniq().cpu(), u)
    assert torch.equal(segs.inverse().cpu()[:nv].long(), inv)
    assert torch.equal(segs.counts().cpu().long(), cnt)
    G = torch.randn(n, 8, generator=g)
    G[nv:] = float("nan")
    dense = torch.zeros(rows, 8).cuda()
    ops.segment_update(segs, ops.RS_UPD_GRAD, 8, 1, dense=G.cuda(), dense_grad=dense)
    want = torch.zeros(rows, 8).index_add_(0, ids[:nv], G[:nv])
    np.testing.assert_allclose(dense.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5)
    ops.check_status()


def test_owner_major_keys():
    """rs_dedup_sort_ex(shard_world): keys = (g % N) * R + g // N, so the distinct rows come out grouped by owner"""
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    cards, N = [10, 1000, 33], 4
    offs = [0, 10, 1010]
    total = sum(cards)
    R = (total + N - 1) // N
    ids = torch.stack([torch.randint(0, c, (500,), generator=g) for c in cards], dim=1)
    segs = ops.dedup_sort(ids.cuda(), 3, offs, total, shard=(N, R))
    glob = (ids + torch.tensor(offs)).reshape(-1)
    keys = (glob % N) * R + glob // N
    u, inv = torch.unique(keys, sorted=True, return_inverse=True)
    assert torch.equal(segs.uniq().cpu(), u) and torch.equal(segs.inverse().cpu().long(), inv)


def _sharded_run(world, kind, D, B, steps, lr, cards, direct_threshold=None, replicate_below=0, shared_exchange=False):
    from deeplearningrecommendationsystem_b200 import dist as rsdist, ops
    from deeplearningrecommendationsystem_b200.nfield import FieldFFM, FieldFM
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    cls = FieldFM if kind == "fm" else FieldFFM
    ref = cls(cards, D, fused=True, seed=4, device="cpu")

    def rank_fn(fab):
        kw = {} if direct_threshold is None else {"direct_threshold": direct_threshold}
        kw["replicate_below"] = replicate_below
        m = cls(cards, D, fused=True, seed=4, device="cuda", sharded=True, fabric=fab, **kw)
        if direct_threshold:
            assert m.direct_fields, "expected some fields to be read directly from the peers' shards"
        assert m.hybrid == bool(replicate_below)
        m.load_global(ref.weight.data)
        tr = Trainer(m, torch.nn.BCELoss(), FusedRowOptimizer(m, torch.optim.SGD([m.bias], lr=lr), lr=lr))
        ids, y = _batch(fab.rank, B, cards)
        ids, y = ids.cuda(), y.cuda()
        preds = []
        for _ in range(steps):
            tr.train_loop(ids, train_rating=y)
            preds.append(tr.predictions_train.detach().clone())
        with torch.no_grad():
            ev = m(ids).clone()                       # inference path: plan + fetch only
        return m.weight.data.clone(), m.bias.detach().clone(), torch.stack(preds), ev, m

    outs = rsdist.ThreadFabric.run(world, rank_fn)
    ops.check_status()
    full = outs[0][4].assemble_global([o[0] for o in outs])
    if replicate_below:     # the replicas must be bit-identical
        for r in range(1, world):
            assert torch.equal(outs[r][4].weight_small.data, outs[0][4].weight_small.data)
    m = cls(cards, D, fused=True, seed=4, device="cpu").cuda()
    tr = Trainer(m, torch.nn.BCELoss(), FusedRowOptimizer(m, torch.optim.SGD([m.bias], lr=lr), lr=lr))
    batches = [_batch(r, B, cards) for r in range(world)]
    ids = torch.cat([b[0] for b in batches]).cuda()
    y = torch.cat([b[1] for b in batches]).cuda()
    for s in range(steps):
        tr.train_loop(ids, train_rating=y)
        want = tr.predictions_train.detach().cpu().numpy()
        for r in range(world):
            np.testing.assert_allclose(outs[r][2][s].cpu().numpy(), want[r * B:(r + 1) * B], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(full.cpu().numpy(), m.weight.detach().cpu().numpy(), rtol=1e-5, atol=2e-6)
    for r in range(world):
        np.testing.assert_allclose(outs[r][1].cpu().numpy(), m.bias.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
    with torch.no_grad():
        ev = m(ids).cpu().numpy()
    for r in range(world):
        np.testing.assert_allclose(outs[r][3].cpu().numpy(), ev[r * B:(r + 1) * B], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("world", [1, 2, 4, 7])
@pytest.mark.parametrize("kind,D", [("fm", 16), ("ffm", 8)])
def test_virtual_ranks_equal_single_gpu(world, kind, D):
    _sharded_run(world, kind, D, 300, 3, 0.5, CARDS)


@pytest.mark.parametrize("world", [1, 3, 4])
def test_virtual_ranks_direct_fields(world):
    """FieldFFM with its big fields read straight from the owners' shards inside the forward kernel (rs_ffm_fwd_peer; the
    exchange flags those requests and rs_shard_serve skips them) == the unsharded step.  Fields of >= 100 rows are direct."""
    _sharded_run(world, "ffm", 8, 300, 3, 0.5, CARDS, direct_threshold=100)
    _sharded_run(world, "ffm", 16, 300, 2, 0.5, CARDS, direct_threshold=100)     # 512-byte rows: the TMA serve kernel with skipped runs


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("kind,D", [("fm", 16), ("ffm", 8), ("ffm", 16)])
def test_virtual_ranks_hybrid_placement(world, kind, D):
    """small tables replicated (dense gradient reduced over the ranks + SGD in rs_replica_sgd), tables of >= 100 rows
    row-sharded (FFM: read from the owners' shards inside the forward kernel, plan built on a side stream; FM: fetched
    into a block) == the unsharded step on the concatenated batch"""
    _sharded_run(world, kind, D, 300, 3, 0.5, CARDS, replicate_below=100)


def test_virtual_ranks_c2_shape_hybrid():
    cards = [min(c, 2000) for c in [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
                                    5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]]
    _sharded_run(4, "ffm", 16, 700, 2, 0.5, cards, replicate_below=2000)
    _sharded_run(4, "fm", 16, 700, 2, 0.5, cards, replicate_below=2000)


def test_virtual_ranks_c2_shape_direct():
    cards = [min(c, 2000) for c in [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
                                    5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]]
    _sharded_run(4, "ffm", 16, 700, 2, 0.5, cards, direct_threshold=2000)


def test_virtual_ranks_c2_shape():
    """F = 26, D = 16: FFM rows of 1664 B take the TMA serve kernel and the streaming segment-reduce with routed stores"""
    cards = [min(c, 2000) for c in [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
                                    5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]]
    _sharded_run(4, "ffm", 16, 700, 2, 0.5, cards)


def test_serve_kernels_agree(monkeypatch):
    """TMA bulk-copy serve kernel == 128-bit load/store serve kernel == torch gather, bit for bit"""
    from deeplearningrecommendationsystem_b200 import dist as rsdist
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    rows, W, n = 5000, 416, 3000
    table = torch.randn(rows, W, generator=g).cuda()
    ids = torch.randint(0, rows, (n, 1), generator=g).cuda()
    blocks = []
    for mode in ("tma", "st"):
        monkeypatch.setenv("RS_SERVE", mode)
        ex = rsdist.DeviceRowExchange(rsdist.LocalFabric())
        plan = ex.plan_for(ids, [0], rows)
        block = ex.fetch(plan, table)
        got = block[plan.local_ids]
        assert torch.equal(got, table[ids.view(-1)]), mode
        blocks.append(block.clone())
    nu = torch.unique(ids).numel()
    assert torch.equal(blocks[0][:nu], blocks[1][:nu])
    ops.check_status()


def test_receive_capacity_overflow_is_flagged():
    """an owner that is asked for more rows than cap_recv sets status bit 16 -> RuntimeError from check_status"""
    from deeplearningrecommendationsystem_b200 import dist as rsdist
    ops = _ops()
    ops.check_status()

    def rank_fn(fab):
        ex = rsdist.DeviceRowExchange(fab, recv_slack=0.3)
        n = 2000
        ids = (torch.arange(n).view(n, 1) * fab.world).cuda()          # every key is owned by rank 0
        table = torch.zeros(ex.local_rows(n * fab.world), 4).cuda()
        plan = ex.plan_for(ids, [0], n * fab.world)
        ex.fetch(plan, table)
        return True

    rsdist.ThreadFabric.run(2, rank_fn)
    with pytest.raises(RuntimeError, match="receive capacity"):
        ops.check_status()


def test_sharded_step_is_graph_capturable():
    """no host sync in the sharded step: it captures into a CUDA graph and replays to the same result as eager steps"""
    from deeplearningrecommendationsystem_b200 import dist as rsdist
    from deeplearningrecommendationsystem_b200.graph import GraphedTrainStep
    from deeplearningrecommendationsystem_b200.nfield import FieldFFM
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    ms = [FieldFFM(CARDS, 8, seed=2, device="cuda", sharded=True, fabric=rsdist.LocalFabric()) for _ in range(2)]
    ms[1].weight.data.copy_(ms[0].weight.data)
    opts = [FusedRowOptimizer(m, torch.optim.SGD([m.bias], lr=0.1), lr=0.1) for m in ms]
    gs = GraphedTrainStep(ms[0], torch.nn.BCELoss(), opts[0], warmup=1)
    tr = Trainer(ms[1], torch.nn.BCELoss(), opts[1])
    for k in range(5):
        ids, y = _batch(k, 256)
        ids, y = ids.cuda(), y.cuda()
        gs(ids, rating=y)
        tr.train_loop(ids, train_rating=y)
    torch.cuda.synchronize()
    assert torch.equal(ms[0].weight.data, ms[1].weight.data) and torch.equal(ms[0].bias.data, ms[1].bias.data)
