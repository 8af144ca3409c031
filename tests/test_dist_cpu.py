"""N > 1 host logic on CPU: world_size-2 gloo processes (the hot path itself has no collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nsa_vibe_b200 import dist as nd


def test_shard_batch_covers_everything_once():
    for B in (0, 1, 2, 7, 8, 64, 513):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                first, n = nd.shard_batch(B, r, world)
                seen += list(range(first, first + n))
            assert seen == list(range(B))
            counts = [nd.shard_batch(B, r, world)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        nd.shard_batch(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # timing reduction = max over ranks; throughput = units of all ranks / that time
        ms = nd.max_over_ranks(10.0 + 5.0 * rank)
        first, n = nd.shard_batch(5, rank, world)
        total = nd.sum_over_ranks(n)
        # batch-sharded "hot path": every rank computes its own rows of a deterministic function, no collective
        x = torch.arange(5 * 3, dtype=torch.float32).reshape(5, 3)
        mine = (x[first:first + n] * 2).sum(dim=1)
        gathered = [torch.zeros(3) for _ in range(world)]
        pad = torch.zeros(3)
        pad[:n] = mine
        dist.all_gather(gathered, pad)  # test-only: collect results to compare with the unsharded run
        # DDP with the reference's gradient hook shape (fp32 allreduce on gloo; bf16 compression needs NCCL)
        model = torch.nn.Linear(3, 2, bias=False)
        with torch.no_grad():
            model.weight.fill_(0.5)
        ddp = torch.nn.parallel.DistributedDataParallel(model)
        ddp(x[first:first + n]).sum().backward()
        # the graph-capturable flat exchange (nd.allreduce_grads_bf16) against DDP's averaged gradient
        m2 = torch.nn.Linear(3, 2, bias=False)
        with torch.no_grad():
            m2.weight.fill_(0.5 + rank)  # differs per rank until broadcast
        nd.broadcast_parameters(m2, 0)
        m2(x[first:first + n]).sum().backward()
        nbytes = nd.allreduce_grads_bf16(m2.parameters(), compress_dtype=torch.float32)  # gloo: no bf16 sum everywhere
        out[rank] = dict(ms=ms, total=total, rows=[g.tolist() for g in gathered], n=n, grad=model.weight.grad.clone(),
                         grad2=m2.weight.grad.clone(), w2=m2.weight.detach().clone(), nbytes=nbytes)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharding_and_max_timing():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0]["ms"] == out[1]["ms"] == 15.0          # max over ranks
    assert out[0]["total"] == out[1]["total"] == 5.0     # all units counted once
    x = torch.arange(15, dtype=torch.float32).reshape(5, 3)
    want = (x * 2).sum(dim=1)
    got = torch.tensor(out[0]["rows"][0][:out[0]["n"]] + out[0]["rows"][1][:out[1]["n"]])
    assert torch.equal(got, want)
    # DDP averaged the per-rank gradients: mean over ranks of sum over that rank's rows
    g0 = x[:3].sum(0).expand(2, 3)
    g1 = x[3:].sum(0).expand(2, 3)
    assert torch.allclose(out[0]["grad"], (g0 + g1) / 2)
    assert torch.allclose(out[1]["grad"], out[0]["grad"])
    for r in (0, 1):
        assert torch.allclose(out[r]["grad2"], out[0]["grad"]) and out[r]["nbytes"] == 12
        assert torch.equal(out[r]["w2"], torch.full((2, 3), 0.5))
