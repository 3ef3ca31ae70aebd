"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sharding arithmetic, DistContext barrier /
all-reduce plumbing, sharded model construction and gather_tables round trip."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from satrans_b200 import SATrans, SparseFeat
from satrans_b200.dist import DistContext, local_rows, shard_of


def test_sharding_arithmetic():
    total = 6571961
    for world in (1, 2, 4, 8):
        assert sum(local_rows(total, r, world) for r in range(world)) == total
        for g in (0, 1, 7, 12345, total - 1):
            r, l = shard_of(g, world)
            assert r == g % world and l * world + r == g and l < local_rows(total, r, world)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ctx = DistContext()
        assert (ctx.rank, ctx.world) == (rank, world)
        t = torch.full((4,), float(rank + 1))
        ctx.all_reduce_sum(t)
        assert torch.equal(t, torch.full((4,), 3.0))
        ctx.barrier('cpu')
        cols = [SparseFeat('a', 11, 32), SparseFeat('b', 6, 32), SparseFeat('301', 5, 32)]
        model = SATrans(cols, cols, ['301'], [3], att_head_num=2, seed=7, device='cpu', flag='sota',
                        table_shard=(rank, world))
        assert model._tables.shape[0] == local_rows(22, rank, world)
        # every rank draws its rows from its own stream: global rows kR .. kR+R-1 are i.i.d., not copies of each other
        fresh = model.gather_tables()
        assert all(not torch.equal(fresh[i], fresh[i + 1]) for i in range(21))
        full = torch.arange(22 * 32, dtype=torch.float32).view(22, 32)
        model.load_full_tables(full)
        assert torch.equal(model._tables, full[rank::world])
        back = model.gather_tables()
        assert torch.equal(back, full)
        ptrs = ctx.exchange_pointers(model._tables)
        assert ptrs[rank] == model._tables.data_ptr() and len(ptrs) == world
        q.put((rank, 'ok'))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res
