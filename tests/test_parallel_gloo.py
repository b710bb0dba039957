"""world_size-2 gloo test (CPU) of the multi-GPU host logic: weight broadcast, utterance sharding, result
gather, max-over-ranks timing reduction."""
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
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from storm_b200 import BackboneRegistry, parallel
        m = BackboneRegistry.get_by_name("ncsnpp")(nf=8, seed=100 + rank, init_scale=1.0)   # ranks start with DIFFERENT weights
        before = parallel.flat_params(m).clone()
        parallel.broadcast_weights(m, src=0)
        after = parallel.flat_params(m)
        ref = parallel.flat_params(BackboneRegistry.get_by_name("ncsnpp")(nf=8, seed=100, init_scale=1.0))
        ok_bcast = torch.equal(after, ref) and (rank == 0 or not torch.equal(before, after)) and m._dirty
        total, L = 7, 5
        ids = parallel.shard_indices(total, rank, world)
        local = torch.stack([torch.full((L,), float(u)) for u in ids])
        gathered = parallel.gather_utterances(local, ids, total, dst=0)
        ok_gather = True
        if rank == 0:
            ok_gather = gathered.shape == (total, L) and all(float(gathered[u, 0]) == u for u in range(total))
        else:
            ok_gather = gathered is None
        mx = parallel.max_over_ranks(10.0 + rank)
        q.put((rank, ok_bcast, ok_gather, mx, ids))
    finally:
        dist.destroy_process_group()


def test_two_rank_plumbing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[4] for r in res] == [[0, 2, 4, 6], [1, 3, 5]]
    for rank, ok_b, ok_g, mx, _ in res:
        assert ok_b and ok_g and mx == 11.0


def test_shard_helper():
    from storm_b200 import parallel
    assert parallel.shard_indices(10, 3, 4) == [3, 7]
    assert sorted(sum((parallel.shard_indices(64, r, 8) for r in range(8)), [])) == list(range(64))
    assert parallel.max_over_ranks(3.0) == 3.0
