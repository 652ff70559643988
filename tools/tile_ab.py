#!/usr/bin/env python
"""A/B of the TMA seed tile (HALGPU_SEED_TILE=1: cp.async.bulk of every interval's run of source top records into shared
memory, liftover_kernel.cuh) against plain loads, on the warp-per-interval walk:
  * C2-faithful with HALGPU_NO_FAST (every interval walked piece by piece: the round-1 regime), sorted and unsorted batch
  * C2-divergent (the walk is the product path there), sorted and unsorted batch
Prints one JSON object; run it under ncu with
  --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,dram__bytes_read.sum
for the stall figures (profiles/r02_tile_ab_*.csv)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import hal_b200
    n = int(os.environ.get("TILE_AB_INTERVALS", "10000000"))
    reps = int(os.environ.get("TILE_AB_REPS", "3"))
    W = bench.WORKLOADS["C2"]
    glen = W["segs"] * bench.SEG_LEN
    gs, ge = bench.make_intervals(n, glen, 2)
    d_gs, d_ge = torch.from_numpy(gs).cuda(), torch.from_numpy(ge).cuda()
    out = {}
    for variant, branch in (("faithful_nofast", "0"), ("divergent", "0.05")):
        hal = bench.ensure_hal("C2", W["segs"], branch)
        with hal_b200.Alignment(hal) as a:
            s, t = a.genome_id(W["src"]), a.genome_id(W["tgt"])
            for order, oflag in (("sorted", 0), ("unsorted", hal_b200.HALGPU_NO_SORT)):
                for tile in ("0", "1"):
                    os.environ["HALGPU_SEED_TILE"] = tile
                    flags = oflag | (hal_b200.HALGPU_NO_FAST if variant == "faithful_nofast" else 0)
                    ms = []
                    for i in range(reps + 1):
                        r = a.liftover_ptrs(s, t, n, d_gs.data_ptr(), d_ge.data_ptr(), None, flags, device=True)
                        if i > 0:
                            ms.append(r.kernel_ms)
                        nrec = r.n_rec
                        r.close()
                    out[f"{variant}/{order}/tile{tile}"] = {"kernel_ms": min(ms), "all": [round(x, 3) for x in ms], "lines": nrec}
    os.environ.pop("HALGPU_SEED_TILE", None)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
