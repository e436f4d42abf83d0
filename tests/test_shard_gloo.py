"""world_size-2 gloo test of the only multi-GPU exchange on the path: region
shard planning, the all_gather of per-shard column summaries and the variable-size payload gather (CPU tensors)."""
import os, socket, sys
import torch
import torch.multiprocessing as mp
from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from samtools_b200 import shard
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    plan = shard.plan_shards(10_000_001, world)
    beg, end = plan[rank]
    # each rank "emits" a number of bytes that depends on its region
    allv, off = shard.gather_summaries([1000 + 7 * rank + (end - beg), end - beg, 3 * rank])
    mx = shard.max_over_ranks(1.5 + rank)
    # variable-size payload gather to rank 0 in shard order (rank r emits 5 + 3 r bytes, all equal to 65 + r)
    n = 5 + 3 * rank
    sizes, _ = shard.gather_summaries([n])
    cat = shard.gather_payload(torch.full((n + 2,), 65 + rank, dtype=torch.uint8), sizes[:, 0].tolist())
    q.put((rank, plan, allv.tolist(), off, mx, None if cat is None else bytes(cat.tolist())))
    dist.destroy_process_group()


def test_region_shards_and_summary_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    plan = res[0][1]
    assert plan[0][0] == 0 and plan[-1][1] == 10_000_001 and plan[0][1] == plan[1][0]     # contiguous, covering
    assert all((e - b) % 4096 == 0 for b, e in plan[:-1])
    for rank, _, allv, off, mx, cat in res:
        assert allv == res[0][2]                      # every rank sees the same table
        assert off == sum(r[0] for r in allv[:rank])   # exclusive prefix of bytes = this shard's output offset
        assert mx == 2.5
        assert cat == (b'A' * 5 + b'B' * 8 if rank == 0 else None)   # payload concatenated in shard order on rank 0 only


def test_single_process_paths():
    sys.path.insert(0, ROOT)
    from samtools_b200 import shard
    allv, off = shard.gather_summaries([5, 6, 7])
    assert allv.tolist() == [[5, 6, 7]] and off == 0
    assert shard.max_over_ranks(3.25) == 3.25
    assert shard.plan_shards(100, 1) == [(0, 100)]
    t = torch.arange(4, dtype=torch.uint8)
    assert shard.gather_payload(t, [4]) is t
