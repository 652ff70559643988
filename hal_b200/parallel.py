"""Multi-GPU plumbing for the liftover path: the staged index is replicated on every GPU,
the interval batch is cut into contiguous shards, one per rank, and the fixed-width output records are brought
together with one all-gather (records padded to the largest per-rank count).  No collective sits inside the walk.

Column sweeps (halAlignmentDepth, BASELINE.json configs[4]) shard the same way: contiguous windows of reference positions,
one per rank, and one all-gather of the fixed-size per-column values.

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


def all_gather_records(counts_per_interval, recs_u8, foreign=True):
    """counts_per_interval: integer tensor (this rank's shard, lines per interval); recs_u8: uint8 tensor of
    sum(counts) * 32 bytes.  Returns (offsets int64 [N+1] for the WHOLE batch in input order, recs uint8) on every rank.
    One small all-gather of the shard sizes, one of the counts, one of the records; equal-sized shards (the weak-scaling
    bench, and any batch of one-line-per-interval lifts) are gathered straight into the result without padding."""
    world = dist.get_world_size()
    dev = recs_u8.device
    counts = counts_per_interval.to(torch.int32)
    if foreign and dev.type == "cuda":
        # The records usually live in memory torch does not own (the library's result buffers).  NCCL is fast on the
        # caching allocator's segments but registers unknown buffers per call (measured: +50 ms per step on 8 GPUs), and
        # the caller wants to free its buffers right after this returns: take one device-to-device copy (~0.2 ms for
        # 320 MB) and wait for it.
        recs_u8 = recs_u8.clone()
        torch.cuda.current_stream(dev).synchronize()
    n_iv, n_rec = counts.numel(), recs_u8.numel() // REC_BYTES
    meta = torch.tensor([n_iv, n_rec], dtype=torch.int64, device=dev)
    metas = torch.empty(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas, meta)
    metas = metas.view(world, 2).tolist()
    max_iv = max(m[0] for m in metas)
    max_rec = max(m[1] for m in metas)
    uniform = all(m[0] == max_iv and m[1] == max_rec for m in metas)
    call = torch.empty(world * max_iv, dtype=torch.int32, device=dev)
    rall = torch.empty(world * max_rec * REC_BYTES, dtype=torch.uint8, device=dev)
    if uniform:
        dist.all_gather_into_tensor(call, counts)
        dist.all_gather_into_tensor(rall, recs_u8)
        all_counts, recs = call, rall
    else:
        cpad = torch.zeros(max_iv, dtype=torch.int32, device=dev)
        cpad[:n_iv] = counts
        rpad = torch.zeros(max_rec * REC_BYTES, dtype=torch.uint8, device=dev)
        rpad[: recs_u8.numel()] = recs_u8
        dist.all_gather_into_tensor(call, cpad)
        dist.all_gather_into_tensor(rall, rpad)
        all_counts = torch.cat([call[r * max_iv: r * max_iv + metas[r][0]] for r in range(world)])
        recs = torch.cat([rall[r * max_rec * REC_BYTES: (r * max_rec + metas[r][1]) * REC_BYTES] for r in range(world)])
    offsets = torch.zeros(all_counts.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(all_counts, 0, dtype=torch.int64, out=offsets[1:])
    return offsets, recs


def all_gather_columns(values, n_total):
    """values: this rank's per-column tensor (its window of shard_bounds(n_total, world), in rank order).  Returns the
    whole sweep (n_total values) on every rank with ONE all-gather: windows differ by at most one column, so every rank
    pads to the largest window and the (at most world - 1) pad slots are dropped afterwards."""
    world = dist.get_world_size()
    bounds = shard_bounds(n_total, world)
    width = max(hi - lo for lo, hi in bounds)
    send = values
    if values.numel() != width:
        send = torch.zeros(width, dtype=values.dtype, device=values.device)
        send[: values.numel()] = values
    out = torch.empty(world * width, dtype=values.dtype, device=values.device)
    dist.all_gather_into_tensor(out, send.contiguous())
    if all(hi - lo == width for lo, hi in bounds):
        return out
    return torch.cat([out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(bounds)])
