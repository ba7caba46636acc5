"""GPU, world_size 2 over NCCL (skipped on a box with fewer than two GPUs): the product's multi-GPU path
`molnextr_b200.parallel.predict_sharded` -- contiguous shards, one engine per rank, one gather of fixed-shape results --
returns on every rank exactly the concatenation of per-shard single-GPU runs (the reference's multi-GPU evaluation has
the same per-shard definition: `DistributedSampler` + `all_gather_object`, main.py:260-302,440-443)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    from molnextr_b200 import synth
    from molnextr_b200.engine import Engine
    from molnextr_b200.parallel import predict_sharded, shard_bounds
    from tests.helpers import seeded_images
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ck = synth.synthetic_checkpoint(0, "sensitised")
    eng = Engine(ck, device=rank, max_batch=4)
    images = seeded_images(321, n, 384, 384)
    out = predict_sharded(eng, images)
    ok = True
    for r in range(world):          # every rank re-computes every shard on its own GPU and compares with the gathered result
        lo, hi = shard_bounds(n, world, r)
        if hi == lo:
            continue
        ref = eng.predict(images[lo:hi].to(eng.device))
        for k in ("ids", "lens", "n_atoms"):
            ok = ok and torch.equal(out[k][lo:hi], ref[k])
        for i, na in enumerate(ref["n_atoms"].tolist()):
            ok = ok and torch.equal(out["edges"][lo + i][:na, :na], ref["edges"][i][:na, :na])
    ok = ok and out["ids"].shape[0] == n
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


@pytest.mark.parametrize("n", [7, 1])
def test_predict_sharded_two_gpus_nccl(n):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    assert res == [(0, True), (1, True)]
