"""Region sharding across GPUs.

The pileup path shards by genome region with no data-path exchange: shard s
owns columns [beg_s, end_s) and stages every read overlapping them (the same
rule as `-r`).  The only collective is the gather of per-shard column
summaries (bytes emitted, columns, reads) to every rank / rank 0, which is what
the emitter needs to concatenate shard outputs in genome order.
Works over NCCL (CUDA tensors) and gloo (CPU tensors; used by the CPU tests).
"""
import torch
import torch.distributed as dist


def plan_shards(length, world, align=4096):
    """Equal-width contiguous shards [beg,end) of a contig, aligned to `align` columns."""
    per = -(-length // world)
    per = -(-per // align) * align
    return [(min(r * per, length), min((r + 1) * per, length)) for r in range(world)]


def gather_summaries(local, device='cpu'):
    """local: sequence of ints (e.g. [bytes_out, n_cols, n_reads]).  Returns a
    (world, len(local)) int64 tensor on every rank (all_gather) plus the
    exclusive prefix of column 0 (this rank's byte offset in the concatenated output)."""
    t = torch.tensor(list(local), dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        allv = torch.stack(out)
        rank = dist.get_rank()
    else:
        allv = t[None, :]
        rank = 0
    offset = int(allv[:rank, 0].sum().item())
    return allv, offset


def max_over_ranks(x, device='cpu'):
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_payload(local, sizes, dst=0):
    """Variable-size gather of the shards' output bytes to rank `dst`, in shard (= genome) order
    (all_gather of byte counts -> exclusive scan -> point-to-point payload).

    local: 1-D uint8 tensor holding this rank's bytes (CUDA tensor over NCCL, CPU tensor over gloo);
    sizes: per-rank byte counts, column 0 of gather_summaries().  Returns the concatenation on `dst`,
    None elsewhere.  Not on bench.py's timed path: a production run lets every rank emit its own
    region (the summaries give the order and the offsets), because funnelling all text through one
    rank's PCIe link would serialise the job; this is for callers that need a single stream."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [int(x) for x in sizes]
    if rank != dst:
        if sizes[rank]:
            dist.send(local[:sizes[rank]].contiguous(), dst=dst)
        return None
    out = torch.empty(sum(sizes), dtype=torch.uint8, device=local.device)
    off = 0
    for r in range(world):
        n = sizes[r]
        if n:
            if r == dst:
                out[off:off + n] = local[:n]
            else:
                dist.recv(out[off:off + n], src=r)
        off += n
    return out


def select_window(soa, beg, end):
    """The reads a region shard [beg,end) stages: every read with pos < end and end position > beg (the `-r` rule,
    bam_plcmd.c:550-554 / htslib's region iterator; zero-span reads count as one base), re-packed as their own batch."""
    import numpy as np
    from . import synth
    pos = soa['pos']
    rl = np.maximum(synth.ref_span(soa), 1)
    idx = np.nonzero((pos < end) & (pos + rl > beg))[0]
    return synth.take_reads(soa, idx)
