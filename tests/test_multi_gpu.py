"""GPU tier (-m gpu): the C++ multi-GPU entry points over real NCCL.  With one GPU the communicator has one rank (NCCL is
still loaded and its all-gather runs); with two or more, one host thread per GPU, each with its own context."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import random_intervals

pytestmark = pytest.mark.gpu


def _gather(world, shards, hal, src, tgt, batches=2):
    import torch
    import hal_b200
    lib = hal_b200.load_library()
    uid = hal_b200.Comm.unique_id(lib)
    out, errs = [None] * world, []

    def rank_main(r):
        try:
            a = hal_b200.Alignment(hal, device=r)
            cm = hal_b200.Comm(a, world, r, uid)
            s, t = a.genome_id(src), a.genome_id(tgt)
            gs, ge, st = shards[r]
            with torch.cuda.device(r):
                dg, de, ds = torch.from_numpy(gs).cuda(), torch.from_numpy(ge).cuda(), torch.from_numpy(st).cuda()
                torch.cuda.synchronize()
                hs = [cm.begin(s, t, len(gs), dg.data_ptr(), de.data_ptr(), ds.data_ptr()) for _ in range(batches)]
                got = []
                for h in hs:
                    res, npr, nrr = cm.end(h)
                    off = np.zeros(res.n + 1, np.uint64)
                    recs = np.zeros(res.n_rec, hal_b200.REC_DTYPE)
                    rt = C.CDLL("libcudart.so.12")
                    assert rt.cudaMemcpy(C.c_void_p(off.ctypes.data), C.c_void_p(res.offsets_ptr), C.c_size_t(off.nbytes), 2) == 0
                    if res.n_rec:
                        assert rt.cudaMemcpy(C.c_void_p(recs.ctypes.data), C.c_void_p(res.recs_ptr), C.c_size_t(recs.nbytes), 2) == 0
                    got.append((off, recs, npr, nrr))
                    res.close()
            out[r] = got
            cm.close()
            a.close()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=600)
    assert not errs, errs
    return out


@pytest.mark.parametrize("wire", ["default", "HALGPU_GATHER_WIRE16", "HALGPU_GATHER_WIRE32", "HALGPU_GATHER_NCCL"])
@pytest.mark.parametrize("ragged", [False, True])
def test_nccl_allgather_equals_single_lift(monkeypatch, ragged, wire):
    import torch
    import hal_b200
    if wire != "default":
        monkeypatch.setenv(wire, "1")
    world = min(torch.cuda.device_count(), 4)
    hal = os.path.join(GOLDEN, "varlen8.hal")
    with hal_b200.Alignment(hal) as a:
        s, t = a.genome_id("L0"), a.genome_id("L3")
        sizes = [5000 + (1000 * r if ragged else 0) for r in range(world)]
        shards = [random_intervals(a.genome_length(s), n, 12 if not ragged else 150, seed=11 + r) for r, n in enumerate(sizes)]
        gs, ge, st = (np.concatenate([sh[k] for sh in shards]) for k in range(3))
        off, recs, _ = a.liftover(s, t, gs, ge, st)
    out = _gather(world, shards, hal, "L0", "L3")
    for r in range(world):
        for goff, grecs, npr, nrr in out[r]:
            assert npr == sizes
            assert np.array_equal(goff, off) and np.array_equal(grecs, recs), f"rank {r}"
