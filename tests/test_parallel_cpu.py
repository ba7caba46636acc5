"""CPU, world_size 2 over gloo: the sharding + gather logic of the multi-GPU path (the engine call is
replaced by a deterministic stand-in; the collective and ordering logic is what is under test)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from molnextr_b200.parallel import gather_predictions, shard_bounds


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
