"""Host-side pieces of bench.py that need no GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_secondary_isolates_failures():
    line = {}
    with bench.Secondary(line, "secondary_x"):
        line["secondary_x"] = {"value": 1.0}
        raise AssertionError("boom")
    assert line["secondary_x"]["value"] == 1.0 and "boom" in line["secondary_x"]["error"]
    with bench.Secondary(line, "secondary_y"):
        raise OSError("nope")
    assert "OSError" in line["secondary_y"]["error"]
    with bench.Secondary(line, "secondary_z"):
        line["secondary_z"] = {"value": 2.0}
    assert line["secondary_z"] == {"value": 2.0}
    with pytest.raises(KeyboardInterrupt):  # only Exceptions are swallowed
        with bench.Secondary(line, "k"):
            raise KeyboardInterrupt


def test_intervals_are_deterministic_and_inside_the_aligned_part():
    glen = 1000 * bench.SEG_LEN
    a = bench.make_intervals(5000, glen, 2)
    b = bench.make_intervals(5000, glen, 2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    gs, ge = a
    assert gs.min() >= 0 and (ge >= gs).all() and ge.max() < glen - bench.SEG_LEN
    ln = ge - gs + 1
    assert ln.min() >= 50 and ln.max() <= 2000
