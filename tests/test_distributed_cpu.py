"""CPU, world_size 2 over gloo: batch sharding and the per-layer gradient all-reduce used by bench.py --gpus N."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from generative_recommenders_b200.distributed import LayerBucketAllReduce, gather_rows, rebase_offsets, shard_sequences


def test_shard_sequences_partition_and_balance():
    g = torch.Generator().manual_seed(0)
    lengths = torch.randint(0, 8192, (37,), generator=g).tolist()
    for world in (1, 2, 4, 8):
        shards = shard_sequences(lengths, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(lengths)))
        loads = [sum(lengths[i] ** 2 for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths) ** 2  # LPT bound
    off = rebase_offsets(lengths, shards[0])
    assert int(off[0]) == 0 and int(off[-1]) == sum(lengths[i] for i in shards[0])
    full_off = torch.zeros(len(lengths) + 1, dtype=torch.int64)
    full_off[1:] = torch.cumsum(torch.tensor(lengths), 0)
    vals = torch.arange(int(full_off[-1])).view(-1, 1)
    rows = gather_rows(vals, full_off, shards[0])
    assert rows.shape[0] == int(off[-1])
    i0 = shards[0][0]
    assert torch.equal(rows[: lengths[i0], 0], torch.arange(int(full_off[i0]), int(full_off[i0 + 1])))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # same initial weights on all ranks
    layers = torch.nn.ModuleList([torch.nn.Linear(8, 8) for _ in range(3)])
    red = LayerBucketAllReduce(list(layers), world, torch.device("cpu"))
    torch.manual_seed(100 + rank)  # different data per rank
    x = torch.randn(5 + rank, 8)
    h = x
    for l in layers:
        h = torch.tanh(l(h))
    h.square().sum().backward()
    red.wait()
    grads = torch.cat([p.grad.reshape(-1) for p in layers.parameters()])
    gathered = [torch.zeros_like(grads) for _ in range(world)]
    dist.all_gather(gathered, grads)
    if rank == 0:
        out["same"] = all(torch.equal(g, gathered[0]) for g in gathered)
        out["launched"] = red.launched
        out["grads"] = grads.clone()
    dist.destroy_process_group()


def _single_reference():
    torch.manual_seed(0)
    layers = torch.nn.ModuleList([torch.nn.Linear(8, 8) for _ in range(3)])
    total = None
    for rank in range(2):
        for p in layers.parameters():
            p.grad = None
        torch.manual_seed(100 + rank)
        x = torch.randn(5 + rank, 8)
        h = x
        for l in layers:
            h = torch.tanh(l(h))
        h.square().sum().backward()
        g = torch.cat([p.grad.reshape(-1) for p in layers.parameters()])
        total = g if total is None else total + g
    return total / 2


@pytest.mark.timeout(120)
def test_layer_bucket_allreduce_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["same"], "ranks disagree after the all-reduce"
    assert out["launched"] == 3, "one bucket per layer"
    torch.testing.assert_close(out["grads"], _single_reference(), rtol=1e-5, atol=1e-6)


def _worker_partial(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    layers = torch.nn.ModuleList([torch.nn.Linear(8, 8) for _ in range(3)])
    red = LayerBucketAllReduce(list(layers), world, torch.device("cpu"))
    layers[1].bias.requires_grad_(False)  # frozen after construction: its hook never fires
    ok = True
    for it in range(2):
        red.zero_grad()
        torch.manual_seed(100 + rank + 10 * it)
        h = torch.randn(4 + rank, 8)
        for i, l in enumerate(layers):
            if it == 0 and i == 2:
                continue  # iteration 0 does not use the last layer at all
            h = torch.tanh(l(h))
        h.square().sum().backward()
        red.wait()
        grads = torch.cat([p.grad.reshape(-1) for p in layers.parameters() if p.grad is not None])
        gathered = [torch.zeros_like(grads) for _ in range(world)]
        dist.all_gather(gathered, grads)
        ok = ok and all(torch.equal(g, gathered[0]) for g in gathered)
        # p.grad stays a view of the flat bucket across iterations
        ok = ok and all(p.grad.data_ptr() == v.data_ptr() for st in red._layers for p, v in zip(st["params"], st["views"]))
    if rank == 0:
        out["ok"] = ok
        out["launched"] = red.launched
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_allreduce_survives_frozen_and_unused_parameters():
    """A layer whose hook count never completes (frozen parameter, unused branch) is reduced in wait(); counters reset."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_partial, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["ok"], "ranks diverged or the gradient views were lost"
    assert out["launched"] == 6, "every layer is reduced exactly once per iteration"
