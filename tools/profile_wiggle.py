"""Run one wiggle liftover of the bench workload (for `ncu -k regex:liftoverKernel`): L7 -> L0 on the C2 alignment."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import hal_b200  # noqa: E402

segs = 1_562_500
hal = bench.ensure_hal(segs)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
nb = min(nb, segs * bench.SEG_LEN - 64)
with hal_b200.Alignment(hal) as a:
    f = np.arange(0, nb, 2048, dtype=np.int64)
    l = np.minimum(f + 2047, nb - 1)
    v = np.random.default_rng(9).random(nb) * 100.0
    for _ in range(2):
        pos, val, info = a.wiggle_liftover(a.genome_id("L7"), a.genome_id("L0"), f, l, f.copy(), v)
    print(len(pos), info)
