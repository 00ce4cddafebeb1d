"""world_size-2 gloo test of the multi-GPU host logic (query sharding + the final
result gather); runs on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crl_kino_b200.planner.sharding import gather_results, reduce_counters, shard_range, shard_sizes


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 8, 1024, 65536, 65537):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_queries, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_queries, rank, world)
    q = torch.arange(lo, hi)
    local = {"status": (q % 3).to(torch.int32), "n_nodes": (q * 2).to(torch.int32),
             "last_state": torch.stack([q.double() * 0.5 + k for k in range(5)], 1)}
    out = gather_results(local, n_queries)
    tot = reduce_counters([float(hi - lo), float(rank + 1)], "sum")
    mx = reduce_counters([float(rank)], "max")
    if rank == 0:
        ret["status"] = out["status"].numpy()
        ret["n_nodes"] = out["n_nodes"].numpy()
        ret["last_state"] = out["last_state"].numpy()
        ret["tot"] = tot
        ret["mx"] = mx
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world2_gloo():
    n = 11                      # ragged: shards of 6 and 5
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n, ret), nprocs=2, join=True)
    q = np.arange(n)
    assert np.array_equal(ret["status"], q % 3)
    assert np.array_equal(ret["n_nodes"], q * 2)
    assert np.allclose(ret["last_state"][:, 3], q * 0.5 + 3)
    assert ret["tot"] == [11.0, 3.0] and ret["mx"] == [1.0]
