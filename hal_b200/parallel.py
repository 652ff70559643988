"""Multi-GPU plumbing for the liftover path: the staged index is replicated on every GPU,
the interval batch is cut into contiguous shards, one per rank, and the fixed-width output records are brought
together with one all-gather (records padded to the largest per-rank count).  No collective sits inside the walk.

Pure torch.distributed: works with the nccl backend on GPUs and with gloo on CPU (tests/test_parallel.py)."""
import torch
import torch.distributed as dist

REC_BYTES = 32


def shard_bounds(n, world):
    """Contiguous, balanced [lo, hi) per rank; concatenating the shards in rank order restores the batch."""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def all_gather_records(counts_per_interval, recs_u8):
    """counts_per_interval: int64 tensor (this rank's shard, lines per interval); recs_u8: uint8 tensor of
    len(sum(counts)) * 32 bytes.  Returns (offsets int64 [N+1] for the WHOLE batch in input order, recs uint8) on every rank."""
    world = dist.get_world_size()
    dev = recs_u8.device
    meta = torch.tensor([counts_per_interval.numel(), recs_u8.numel() // REC_BYTES], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    max_iv = int(max(int(m[0]) for m in metas))
    max_rec = int(max(int(m[1]) for m in metas))
    cpad = torch.zeros(max_iv, dtype=torch.int64, device=dev)
    cpad[: counts_per_interval.numel()] = counts_per_interval
    rpad = torch.zeros(max_rec * REC_BYTES, dtype=torch.uint8, device=dev)
    rpad[: recs_u8.numel()] = recs_u8
    call = torch.empty(world * max_iv, dtype=torch.int64, device=dev)
    rall = torch.empty(world * max_rec * REC_BYTES, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(call, cpad)
    dist.all_gather_into_tensor(rall, rpad)
    counts = torch.cat([call[r * max_iv: r * max_iv + int(metas[r][0])] for r in range(world)])
    recs = torch.cat([rall[r * max_rec * REC_BYTES: (r * max_rec + int(metas[r][1])) * REC_BYTES] for r in range(world)])
    offsets = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(counts, 0)
    return offsets, recs
