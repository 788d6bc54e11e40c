"""N>1 path on CPU: world_size-2 gloo run of the sharding / timing / gather logic used by bench.py under torchrun."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "deep-image-prior_b200"))
    import multi_gpu as mg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = mg.shard(5, rank, world)
    dist.barrier()
    t = mg.max_over_ranks(1.0 + rank)                      # rank 1 is the slow one
    recs = mg.gather_records([20.0 + rank, 0.01 * (rank + 1), 100.0 / (1.0 + rank)])
    q.put((rank, mine, t, recs, mg.aggregate_rate(10, t, world)))
    dist.destroy_process_group()


def test_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]            # static image -> rank sharding, no overlap
    for rank, mine, t, recs, rate in res:
        assert t == 2.0                                              # max over ranks
        assert recs == [[20.0, 0.01, 100.0], [21.0, 0.02, 50.0]]     # every rank sees every record
        assert rate == 2 * 10 / 2.0


def test_single_process_degenerates():
    import multi_gpu as mg
    assert mg.shard(3, 0, 1) == [0, 1, 2]
    assert mg.max_over_ranks(3.5) == 3.5
    assert mg.gather_records([1.0, 2.0]) == [[1.0, 2.0]]
