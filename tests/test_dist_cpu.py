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


def _worker_bf16(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)
        # --- the compressed exchange itself: divide by world, round to bf16, sum, copy back (train_showcase.py:654-665) ---
        g = torch.randn(4096) * (1 + rank)
        p = torch.nn.Parameter(torch.zeros(4096))
        p.grad = g.clone()
        nbytes = nd.allreduce_grads_bf16([p])
        parts = [torch.empty(4096, dtype=torch.bfloat16) for _ in range(world)]
        dist.all_gather(parts, (g / world).to(torch.bfloat16))
        raw = [torch.empty(4096) for _ in range(world)]
        dist.all_gather(raw, g)
        # --- bucketed, overlapped exchange == the flat one, bucket by bucket ---
        def net():
            torch.manual_seed(7)
            return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
        x = torch.randn(8, 6) * (1 + rank)
        a, b = net(), net()
        ex = nd.OverlappedGradExchange([list(a[4].parameters()), list(a[2].parameters()), list(a[0].parameters())])
        a(x).square().sum().backward()
        fired = len(ex._pending)              # buckets launched from inside backward
        nb = ex.finish()
        b(x).square().sum().backward()
        nd.allreduce_grads_bf16(b.parameters())
        # second step: hooks re-arm, a parameter without a gradient does not wedge its bucket
        for q in a.parameters():
            q.grad = None
        a[0].weight.requires_grad_(False)
        a(x).square().sum().backward()
        nb2 = ex.finish()
        out[rank] = dict(got=p.grad.clone(), parts=[t.clone() for t in parts], raw=[t.clone() for t in raw], nbytes=nbytes, fired=fired, nb=nb,
                         ga=[q.grad.clone() for q in b.parameters()], gb=[q.grad.clone() for q in b.parameters()],
                         gx=[q.grad.clone() if q.grad is not None else None for q in a.parameters()], nb2=nb2)
        out[rank]["ga"] = None  # replaced below (a's grads were reset for the second step); keep the comparison on the first step:
    finally:
        dist.destroy_process_group()


def _worker_overlap_equals_flat(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def net():
            torch.manual_seed(7)
            return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
        torch.manual_seed(50 + rank)
        x = torch.randn(8, 6) * (1 + rank)
        a, b = net(), net()
        ex = nd.OverlappedGradExchange([list(a[4].parameters()), list(a[2].parameters()), list(a[0].parameters())])
        a(x).square().sum().backward()
        fired = len(ex._pending)
        nb = ex.finish()
        b(x).square().sum().backward()
        nd.allreduce_grads_bf16(b.parameters())
        out[rank] = dict(fired=fired, nb=nb, ga=[q.grad.clone() for q in a.parameters()], gb=[q.grad.clone() for q in b.parameters()])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_bf16_compressed_exchange_divides_then_rounds(world):
    """the bf16 arithmetic of the exchange was never asserted.  gloo sums bf16 like NCCL does; a world of three makes
    the order of divide and round observable (g/3 is not exact in bf16)."""
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker_bf16, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        o = out[r]
        assert o["nbytes"] == 4096 * 2
        want = torch.stack([t.float() for t in o["parts"]]).sum(0)              # sum over ranks of bf16(g_r / world), in fp32
        # summed in bf16 or better: every hop rounds the running sum once, |error| <= (world - 1) * 2^-8 * sum_r |part_r|
        bound = (world - 1) * 2.0 ** -8 * torch.stack([t.float().abs() for t in o["parts"]]).sum(0)
        assert bool(((o["got"] - want).abs() <= bound + 1e-30).all())
        ulp = bound / (world - 1) * 2
        assert torch.equal(o["got"], out[0]["got"])                             # every rank holds the same reduced gradient
        if world == 3:  # the other order, bf16(g_r) / world summed, is measurably different: the test can tell them apart
            other = torch.stack([t.to(torch.bfloat16).float() / world for t in o["raw"]]).sum(0)
            exact = torch.stack(o["raw"]).sum(0) / world
            assert (want - exact).abs().mean() != (other - exact).abs().mean()
            assert ((o["got"] - want).abs() <= (o["got"] - other).abs() + ulp).float().mean() > 0.9
        # second step of the bucketed exchange: the frozen parameter has no gradient, everything else was reduced
        assert o["gx"][0] is None and all(g is not None for g in o["gx"][1:]) and o["nb2"] > 0


def test_overlapped_bucket_exchange_equals_flat_exchange():
    world = 2
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker_overlap_equals_flat, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        o = out[r]
        assert o["fired"] == 3, "every bucket must be launched from inside backward, not at finish()"
        assert o["nb"] == 2 * sum(g.numel() for g in o["ga"])
        for ga, gb in zip(o["ga"], o["gb"]):
            assert torch.equal(ga, gb)          # same values as the one flat exchange (element-wise sums: bucketing cannot change them)
        for ga, g0 in zip(o["ga"], out[0]["ga"]):
            assert torch.equal(ga, g0)
