"""CPU suite: the N > 1 host logic (pair sharding, weight broadcast, result gather) on world_size-2 gloo."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gtsfm_b200 import distributed as D


def test_shard_pairs_partition_is_exact():
    pairs = [(i, j) for i in range(12) for j in range(i + 1, 12)]
    for world in (1, 2, 3, 8):
        shards = [D.shard_pairs(pairs, r, world) for r in range(world)]
        assert sorted(p for s in shards for p in s) == sorted(pairs)
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
        assert shards[0][:2] == [pairs[0], pairs[world]] if len(pairs) > world else True
    assert D.images_needed([(3, 5), (5, 9)]) == [3, 5, 9]


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        order = ["a.weight", "a.bias", "scalar"]
        truth = {"a.weight": np.arange(12, dtype=np.float32).reshape(3, 4), "a.bias": np.ones(3, np.float32), "scalar": np.array(2.5, np.float32)}
        sd = truth if rank == 0 else {k: np.zeros_like(v) for k, v in truth.items()}
        got = D.broadcast_state_dict(sd, order, src=0)
        ok_bcast = all(np.array_equal(got[k], truth[k]) and got[k].shape == truth[k].shape for k in order)
        pairs = [(i, i + 1) for i in range(7)]
        mine = D.shard_pairs(pairs, rank, world)
        local = {p: np.full((p[0] + 1, 2), rank, np.int64) for p in mine}
        merged = D.gather_pair_results(local)
        ok_gather = sorted(merged) == pairs and all(merged[p].shape == (p[0] + 1, 2) and int(merged[p][0, 0]) == (i % world)
                                                    for i, p in enumerate(pairs))
        # the feature exchange of the strong-scaling seam: rank r owns the images at positions r, r + world, ...
        n, k = 3, 5
        kp = torch.full((n, k, 2), float(rank)); sc = torch.full((n, k), 10.0 + rank); de = torch.full((n, k, 256), 20.0 + rank)
        cnt = torch.tensor([k, rank + 1, 0], dtype=torch.int32)
        akp, asc, ade, acnt = D.all_gather_features(kp, sc, de, cnt)
        ok_gather = ok_gather and akp.shape == (world * n, k, 2) and acnt.tolist() == [k, 1, 0, k, 2, 0]
        ok_gather = ok_gather and all(float(akp[r * n + j, 0, 0]) == r and float(asc[r * n + j, 0]) == 10 + r and float(ade[r * n + j, 0, 0]) == 20 + r
                                      for r in range(world) for j in range(n))
        ok_gather = ok_gather and [D.image_owner(i, world) for i in range(5)] == [0, 1, 0, 1, 0]
        # the bench's timing reduction: max over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, ok_bcast, ok_gather, float(t.item())))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True, 11.0), (1, True, True, 11.0)]
