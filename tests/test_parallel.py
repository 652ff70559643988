"""CPU tier: the N>1 path (shard -> lift -> one all-gather of the output records) under gloo with world_size 2.
Each rank lifts its shard with the emulated library; the gathered result must equal the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT
from helpers import random_intervals


def _worker(rank, world, port, emul_lib, hal, out_dir):
    sys.path.insert(0, ROOT)
    import hal_b200
    from hal_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = hal_b200.Alignment(hal, lib_path=emul_lib)
    s, t = a.genome_id("L0"), a.genome_id("L3")
    gs, ge, st = random_intervals(a.genome_length(s), 301, 250, seed=77)
    lo, hi = parallel.shard_bounds(len(gs), world)[rank]
    off, recs, _ = a.liftover(s, t, gs[lo:hi], ge[lo:hi], st[lo:hi])
    counts = torch.from_numpy(np.diff(off.astype(np.int64)))
    offsets, allrecs = parallel.all_gather_records(counts, torch.from_numpy(recs.view(np.uint8).copy()))
    if rank == 0:
        np.save(os.path.join(out_dir, "offsets.npy"), offsets.numpy())
        np.save(os.path.join(out_dir, "recs.npy"), allrecs.numpy())
    a.close()
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    from hal_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 1001):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_two_rank_gloo_equals_single_process(emul_lib, tmp_path):
    import hal_b200
    hal = os.path.join(GOLDEN, "varlen8.hal")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, emul_lib, hal, str(tmp_path)), nprocs=2, join=True)
    a = hal_b200.Alignment(hal, lib_path=emul_lib)
    s, t = a.genome_id("L0"), a.genome_id("L3")
    gs, ge, st = random_intervals(a.genome_length(s), 301, 250, seed=77)
    off, recs, _ = a.liftover(s, t, gs, ge, st)
    a.close()
    assert np.array_equal(np.load(tmp_path / "offsets.npy"), off.astype(np.int64))
    assert np.array_equal(np.load(tmp_path / "recs.npy"), recs.view(np.uint8))


def _depth_worker(rank, world, port, emul_lib, hal, out_dir):
    sys.path.insert(0, ROOT)
    import hal_b200
    from hal_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = hal_b200.Alignment(hal, lib_path=emul_lib)
    g = a.genome_id("L1")
    n = 4001  # odd: the windows differ by one column
    lo, hi = parallel.shard_bounds(n, world)[rank]
    d, _ = a.depth(g, lo, hi - 1, 1, (), hal_b200.HALGPU_COUNT_DUPES)
    whole = parallel.all_gather_columns(torch.from_numpy(d), n)
    if rank == 1:
        np.save(os.path.join(out_dir, "depth.npy"), whole.numpy())
    a.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_depth_sweep_equals_single_process(emul_lib, tmp_path):
    """halAlignmentDepth sweep sharded by reference window + one all-gather == the single-process sweep"""
    import hal_b200
    hal = os.path.join(GOLDEN, "varlen8.hal")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_depth_worker, args=(2, port, emul_lib, hal, str(tmp_path)), nprocs=2, join=True)
    a = hal_b200.Alignment(hal, lib_path=emul_lib)
    d, _ = a.depth(a.genome_id("L1"), 0, 4000, 1, (), hal_b200.HALGPU_COUNT_DUPES)
    a.close()
    assert np.array_equal(np.load(tmp_path / "depth.npy"), d)
