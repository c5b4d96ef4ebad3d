"""CPU test of the N>1 plumbing with the gloo backend, world_size 2: member sharding covers
every member exactly once, the motion-field broadcast delivers rank 0's field, and timings
reduce to the slowest rank."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from pysteps_b200 import _shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = _shard.member_indices(24, world, rank)
    field = torch.arange(2 * 6 * 5, dtype=torch.float64).reshape(2, 6, 5) * 0.25 if rank == 0 else None
    got = _shard.broadcast_field(field, src=0, shape=(2, 6, 5))
    slow = _shard.max_over_ranks(1.0 + rank)
    # row bands of a composite, gathered back into the full field (even and uneven splits)
    gathered = []
    for m in (8, 7):
        whole = torch.arange(2 * m * 5, dtype=torch.float64).reshape(2, m, 5)
        r0, r1 = _shard.row_band(m, world, rank)
        gathered.append(torch.equal(_shard.gather_row_bands(whole[:, r0:r1].contiguous(), m, world, rank), whole))
    q.put((rank, mine, got.numpy().copy(), slow, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_broadcast_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    members = sorted(i for _, mine, _, _, _ in res for i in mine)
    assert members == list(range(24))
    ref = (np.arange(60, dtype=np.float64) * 0.25).reshape(2, 6, 5)
    for rank, mine, got, slow, gathered in res:
        assert gathered == [True, True]
        assert len(mine) == 12
        assert np.array_equal(got, ref)
        assert slow == 2.0


def test_single_process_noop():
    from pysteps_b200 import _shard
    f = torch.ones(2, 3, 3)
    assert _shard.broadcast_field(f) is f
    assert _shard.max_over_ranks(3.5) == 3.5
    assert _shard.member_indices(5, 1, 0) == [0, 1, 2, 3, 4]
    bands = [_shard.row_band(4097, 8, r) for r in range(8)]
    assert bands[0][0] == 0 and bands[-1][1] == 4097
    assert all(bands[i][1] == bands[i + 1][0] for i in range(7))
    assert max(b[1] - b[0] for b in bands) - min(b[1] - b[0] for b in bands) <= 1
