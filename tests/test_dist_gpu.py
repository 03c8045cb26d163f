"""N-GPU row-sharded FM / FFM train steps == the 1-GPU step on the concatenated global batch (needs >= 2 GPUs), for both
exchange implementations: dist.DeviceRowExchange (this package's kernels over NVLink peer memory, the default) and
dist.RowExchange (NCCL all-to-alls).  On a 1-GPU box these skip; tests/test_shard_gpu.py runs the same comparison with
virtual ranks, and `bench.py --gpus N` prints the same check ("sharded_equals_single") in its JSON line."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
CARDS = [3, 50, 7, 1000, 24, 12, 301, 5]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(rank, B):
    g = torch.Generator().manual_seed(500 + rank)
    ids = torch.stack([(torch.rand(B, generator=g) ** 2 * c).long().clamp_(max=c - 1) for c in CARDS], dim=1)
    return ids, (torch.rand(B, 1, generator=g) < 0.3).float()


def _worker(rank, world, port, kind, D, B, steps, lr, out, exchange):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RS_PEER_EXCHANGE="1" if exchange == "device" else "0")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from deeplearningrecommendationsystem_b200 import dist as rsdist, ops
        from deeplearningrecommendationsystem_b200.nfield import FieldFFM, FieldFM
        from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
        from deeplearningrecommendationsystem_b200.trainer import Trainer
        cls = FieldFM if kind == "fm" else FieldFFM
        ref = cls(CARDS, D, fused=True, seed=4, device="cpu")                  # same global init on every rank
        m = cls(CARDS, D, fused=True, seed=4, device=f"cuda:{rank}", sharded=True)
        m.load_global(ref.weight.data)
        tr = Trainer(m, torch.nn.BCELoss(), FusedRowOptimizer(m, torch.optim.SGD([m.bias], lr=lr), lr=lr))
        ids, y = _batch(rank, B)
        preds = []
        for _ in range(steps):
            tr.train_loop(ids.cuda(), train_rating=y.cuda())
            preds.append(tr.predictions_train.detach().cpu())
        ops.check_status()
        shards = [torch.empty((m.total_rows - r + world - 1) // world, m.width, device=f"cuda:{rank}") for r in range(world)]
        # gather every shard on rank 0 through padded all_gather
        maxr = max(s.shape[0] for s in shards)
        pad = torch.zeros(maxr, m.width, device=f"cuda:{rank}")
        pad[: m.weight.shape[0]] = m.weight.data
        got = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(got, pad)
        if rank == 0:
            full = rsdist.unshard_rows([got[r][: shards[r].shape[0]] for r in range(world)])
            torch.save({"table": full.cpu(), "bias": m.bias.detach().cpu(), "pred0": torch.stack(preds)}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("exchange", ["device", "nccl"])
@pytest.mark.parametrize("kind,D", [("fm", 16), ("ffm", 8), ("ffm", 16)])
def test_sharded_equals_single_gpu(kind, D, exchange, world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from deeplearningrecommendationsystem_b200.nfield import FieldFFM, FieldFM
    from deeplearningrecommendationsystem_b200.optim import FusedRowOptimizer
    from deeplearningrecommendationsystem_b200.trainer import Trainer
    B, steps, lr = 300, 3, 0.5
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), kind, D, B, steps, lr, out, exchange), nprocs=world, join=True)
    res = torch.load(out)
    cls = FieldFM if kind == "fm" else FieldFFM
    m = cls(CARDS, D, fused=True, seed=4, device="cpu")
    m = m.cuda()
    tr = Trainer(m, torch.nn.BCELoss(), FusedRowOptimizer(m, torch.optim.SGD([m.bias], lr=lr), lr=lr))
    batches = [_batch(r, B) for r in range(world)]
    ids = torch.cat([b[0] for b in batches]).cuda()
    y = torch.cat([b[1] for b in batches]).cuda()
    for s in range(steps):
        tr.train_loop(ids, train_rating=y)
        np.testing.assert_allclose(res["pred0"][s].numpy(), tr.predictions_train.detach().cpu().numpy()[:B], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(res["table"].numpy(), m.weight.detach().cpu().numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(res["bias"].numpy(), m.bias.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
