"""CPU tier: the C-ABI library builds for sm_100a, loads, exports every symbol include/halgpu.h declares,
and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import GOLDEN, ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "halgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(halgpu_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(product_lib):
    import hal_b200
    lib = ctypes.CDLL(product_lib)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/halgpu.h but not exported"
    assert sorted(hal_b200.ABI_SYMBOLS) == syms


def test_library_contains_sm100a_kernels(product_lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", product_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import hal_b200
    with pytest.raises(hal_b200.HalGpuError):
        hal_b200.Alignment(os.path.join(GOLDEN, "varlen8.hal"))


def test_missing_library_fails_loudly(tmp_path):
    import hal_b200
    with pytest.raises(hal_b200.HalGpuError):
        hal_b200.load_library(str(tmp_path / "nope.so"))
