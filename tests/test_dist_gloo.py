"""World-size-2 CPU test (gloo) of the N>1 host logic: batch sharding and the final all-gather of decoded ids
keep the global order and are bit-identical to the single-rank result."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))


def _worker(rank, world, port, q):
    from ldmseg.engine.dist import shard_range, gather_ids
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gb = 6
    g = torch.Generator().manual_seed(0)
    global_ids = (torch.rand(gb, 8, 8, generator=g) * 128).to(torch.uint8)      # stand-in for decoded ids
    lo, hi = shard_range(gb, rank, world)
    out = gather_ids(global_ids[lo:hi].clone(), gb)
    q.put((rank, lo, hi, torch.equal(out, global_ids)))
    dist.barrier()
    dist.destroy_process_group()


def test_batch_shard_and_gather_gloo():
    from ldmseg.engine.dist import shard_range
    assert [shard_range(64, r, 8) for r in range(8)] == [(8 * r, 8 * r + 8) for r in range(8)]
    assert [shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:3] == (0, 3) and res[1][1:3] == (3, 6)
    assert all(r[3] for r in res)
