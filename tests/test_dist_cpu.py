"""Row-sharded exchange (dist.RowExchange) on CPU: world_size 2 and 3 over gloo, torch-CPU primitives injected.

Checks the routing/index math of the multi-GPU path: every lookup gets exactly its global row back, and the
per-row gradients pushed to the owners reproduce the single-process dense gradient, shard by shard."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_prims():
    from deeplearningrecommendationsystem_b200.dist import Prims

    def unique(keys, total_rows):
        u, inv = torch.unique(keys, sorted=True, return_inverse=True)
        return u, inv

    return Prims(unique, lambda table, idx: table[idx])


def _worker(rank, world, port, total_rows, W, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deeplearningrecommendationsystem_b200 import dist as rsdist
        g = torch.Generator().manual_seed(1)
        table = torch.randn(total_rows, W, generator=g)                   # the global table, same on every rank
        gk = torch.Generator().manual_seed(100 + rank)
        keys = (torch.rand(n, generator=gk) ** 2 * total_rows).long().clamp_(max=total_rows - 1)   # skewed, many duplicates
        G = torch.randn(n, W, generator=gk)                                # per-lookup row gradients of this rank
        ex = rsdist.RowExchange(_cpu_prims())
        assert ex.local_rows(total_rows) == table[rank::world].shape[0]
        shard = rsdist.shard_rows(table, rank, world)
        plan = ex.plan(keys, total_rows)
        block = ex.fetch(plan, shard)
        assert plan.n_uniq == torch.unique(keys).numel() == block.shape[0]
        assert torch.equal(block[plan.local_ids], table[keys])             # forward: bit-exact rows
        # backward: reduce per fetched row, push to owners, owners sum what they receive
        block_grad = torch.zeros(plan.n_uniq, W).index_add_(0, plan.local_ids, G)
        recv = ex.push_grads(plan, block_grad)
        assert recv.shape[0] == plan.recv_local.numel()
        shard_grad = torch.zeros_like(shard).index_add_(0, plan.recv_local, recv)
        # oracle: dense gradient of the global table from ALL ranks' lookups
        all_keys = [torch.empty(n, dtype=torch.int64) for _ in range(world)]
        all_G = [torch.empty(n, W) for _ in range(world)]
        dist.all_gather(all_keys, keys)
        dist.all_gather(all_G, G)
        dense = torch.zeros(total_rows, W).index_add_(0, torch.cat(all_keys), torch.cat(all_G))
        torch.testing.assert_close(shard_grad, dense[rank::world], rtol=1e-5, atol=1e-5)
        # replicated dense parameters: one averaged all-reduce
        p = torch.nn.Parameter(torch.zeros(3))
        p.grad = torch.full((3,), float(rank + 1))
        rsdist.allreduce_dense_grads([p])
        assert torch.allclose(p.grad, torch.full((3,), sum(range(1, world + 1)) / world))
        # round trip of the shard helpers
        shards = [torch.empty_like(table[r::world]) for r in range(world)]
        for r in range(world):
            shards[r] = table[r::world].contiguous()
        assert torch.equal(rsdist.unshard_rows(shards), table)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_exchange_gloo(world):
    mp.spawn(_worker, args=(world, _free_port(), 101, 4, 500), nprocs=world, join=True)


def test_row_exchange_single_process():
    from deeplearningrecommendationsystem_b200 import dist as rsdist
    ex = rsdist.RowExchange(_cpu_prims())
    table = torch.arange(40.0).view(10, 4)
    keys = torch.tensor([3, 3, 9, 0, 3])
    plan = ex.plan(keys, 10)
    block = ex.fetch(plan, table)
    assert plan.n_uniq == 3 and torch.equal(block[plan.local_ids], table[keys])
    assert torch.equal(ex.push_grads(plan, block), block)
