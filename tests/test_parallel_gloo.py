"""N>1 host logic on CPU: world_size-2 gloo processes exercise the env sharding and the episode-statistics all-reduce."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from panda_lang_manip_b200.parallel import all_reduce_stats, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, stop = shard_range(65537, rank, world)
    local = np.array([stop - start, rank + 1.0, -10.0 * (rank + 1), 50.0 * (stop - start)])
    total = all_reduce_stats(local)
    out.put((rank, start, stop, total.tolist()))
    dist.destroy_process_group()


def test_shard_and_stats_all_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    (r0, s0, e0, t0), (r1, s1, e1, t1) = res
    assert (s0, e0, s1, e1) == (0, 32769, 32769, 65537)
    assert t0 == t1 == [65537.0, 3.0, -30.0, 50.0 * 65537]


def test_shard_range_covers_everything():
    from panda_lang_manip_b200.parallel import shard_range
    for total in (1, 7, 8, 262144):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
