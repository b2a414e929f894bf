"""World-size-2 gloo run of the host-side multi-GPU logic: env sharding and the episode-statistics all-reduce
(the only collective of the design)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ccbs_b200 import dist as cd, lib as L
    start, stop = cd.shard_range(65536 + 3, rank, world)
    acc = torch.zeros(L.NUM_ACCUM, dtype=torch.float64)
    acc[0], acc[1], acc[2], acc[3] = 10 + rank, -100.0 * (rank + 1), 250 + rank, rank
    vals = cd.reduce_episode_stats(acc)
    out[rank] = (start, stop, vals["episodes"], vals["ep_rew_mean"], vals["ep_len_mean"], vals["win_rate"])
    dist.destroy_process_group()


def test_shards_and_stat_reduce():
    from ccbs_b200 import dist as cd
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29541, out), nprocs=world, join=True)
    (s0, e0, n0, r0, l0, w0), (s1, e1, n1, r1, l1, w1) = out[0], out[1]
    assert (s0, e1) == (0, 65539) and e0 == s1 and (e0 - s0) - (e1 - s1) in (0, 1)
    assert n0 == n1 == 21 and r0 == r1 == -300.0 / 21 and l0 == l1 == 501 / 21 and w0 == 1 / 21
    # the partition is exact for any world size
    for world in (1, 3, 8):
        edges = [cd.shard_range(1000, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
