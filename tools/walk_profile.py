#!/usr/bin/env python
"""Two lifts of the first N intervals of the bench batch over the DIVERGENT C2 file (every interval takes the warp-per-interval
walk).  For `ncu --set full --import-source on -k regex:liftoverKernel -s 3 -c 1`: the first call's launches are rung 1, the
pool-full re-run and the scratch rung; launch 3 is the second call's rung 1 with the record pool sized right."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import hal_b200
    n = int(os.environ.get("WALK_INTERVALS", "2000000"))
    W = bench.WORKLOADS["C2"]
    gs, ge = bench.make_intervals(10_000_000, W["segs"] * bench.SEG_LEN, 2)
    d_gs, d_ge = torch.from_numpy(gs[:n]).cuda(), torch.from_numpy(ge[:n]).cuda()
    hal = bench.ensure_hal("C2", W["segs"], "0.05")
    with hal_b200.Alignment(hal) as a:
        s, t = a.genome_id(W["src"]), a.genome_id(W["tgt"])
        for _ in range(2):
            r = a.liftover_ptrs(s, t, n, d_gs.data_ptr(), d_ge.data_ptr(), None, 0, device=True)
            print(r.kernel_ms, r.n_rec, r.n_retry, r.launches)
            r.close()


if __name__ == "__main__":
    main()
