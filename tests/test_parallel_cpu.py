"""CPU, world_size 2 over gloo: the sharding + gather logic of the multi-GPU path (the engine call is
replaced by a deterministic stand-in; the collective and ordering logic is what is under test)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from molnextr_b200.parallel import gather_predictions, predict_sharded, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 32, 2048):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _fake_local(lo, hi):
    rows = torch.arange(lo, hi)
    return {"ids": (rows.view(-1, 1) * 10 + torch.arange(5).view(1, -1)).int(), "lens": (rows % 5 + 1).int(),
            "edges": (rows.view(-1, 1, 1) + torch.zeros(1, 3, 3)).to(torch.uint8)}


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n, world, rank)
    out = gather_predictions(_fake_local(lo, hi), n)
    ref = _fake_local(0, n)
    ok = all(torch.equal(out[k], ref[k]) for k in ref)
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 7   # ragged: shards of 4 and 3
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


class _FakeEngine:
    """Stand-in with the attributes predict_sharded uses; `predict` fails loudly on an empty shard like mnx_encode."""
    device = torch.device("cpu")

    def __init__(self, max_batch, lo):
        self.max_batch, self.lo = max_batch, lo

    def predict(self, x):
        assert x.shape[0] > 0, "engine called with an empty shard"
        return _fake_local(self.lo, self.lo + x.shape[0])

    def empty_result(self):
        return _fake_local(0, 0)


def _worker_sharded(rank, world, port, n, max_batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, _ = shard_bounds(n, world, rank)
    try:
        out = predict_sharded(_FakeEngine(max_batch, lo), torch.zeros((n, 3, 4, 4)))
        ref = _fake_local(0, n)
        res = all(torch.equal(out[k], ref[k]) for k in ref)
    except ValueError:
        res = "rejected"
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def _run_sharded(n, max_batch):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, n, max_batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    return res


def test_sharded_fewer_images_than_ranks_gloo():
    """n = 1 on 2 ranks: rank 1 has an empty shard, skips the engine and still enters the all-gather."""
    assert _run_sharded(1, 8) == [(0, True), (1, True)]


def test_sharded_over_capacity_is_rejected_on_every_rank_gloo():
    """shards of 5 and 4 rows against max_batch = 4: BOTH ranks raise before any collective (no hang)."""
    assert _run_sharded(9, 4) == [(0, "rejected"), (1, "rejected")]
