import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_cases():
    return json.load(open(os.path.join(GOLDEN, "cases", "index.json")))


@pytest.fixture(scope="session")
def oracle_lib():
    """Builds oracle/liboracle.so (the CPU restatement) on demand."""
    import pyoracle
    if not os.path.exists(pyoracle.LIB):
        pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def product_lib():
    """hal_b200/libhalgpu.so (built for sm_100a; nvcc cross-compiles without a GPU)."""
    from hal_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def emul_lib():
    """tests/simt/libhalgpu_emul.so: the SAME kernel/engine sources compiled for the host warp emulator."""
    out = os.path.join(ROOT, "tests", "simt", "libhalgpu_emul.so")
    csrc = os.path.join(ROOT, "hal_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in ("capi.cu", "engine.cu", "multi.cu", "halmmap.cpp")]
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc) if os.path.isfile(os.path.join(csrc, f))] + \
           [os.path.join(ROOT, "tests", "simt", "simt_emul.h"), os.path.join(ROOT, "include", "halgpu.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cmd = ["g++", "-std=c++17", "-O2", "-DHALGPU_SIMT_EMUL", "-I" + os.path.join(ROOT, "tests", "simt"), "-I" + csrc,
               "-fPIC", "-shared", "-pthread"]
        for s in srcs:
            cmd += ["-x", "c++", s]
        subprocess.check_call(cmd + ["-o", out])
    return out


@pytest.fixture(scope="session")
def emul_cli(emul_lib):
    """The halLiftover CLI (hal_b200/csrc/host) linked against the emulated library."""
    out = os.path.join(ROOT, "tests", "simt", "halLiftover_emul")
    host = os.path.join(ROOT, "hal_b200", "csrc", "host")
    srcs = [os.path.join(host, f) for f in ("halLiftoverMain.cpp", "gpu_liftover.cpp", "bed.cpp", "bed_fast.cpp")]
    deps = srcs + [os.path.join(host, f) for f in ("gpu_liftover.hpp", "bed.hpp", "bed_fast.hpp")] + [emul_lib]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", out] + srcs +
                              ["-L" + os.path.dirname(emul_lib), "-lhalgpu_emul", "-Wl,-rpath,$ORIGIN", "-pthread"])
    return out


@pytest.fixture(scope="session")
def emul_wig_cli(emul_lib):
    """The halWiggleLiftover CLI linked against the emulated library."""
    out = os.path.join(ROOT, "tests", "simt", "halWiggleLiftover_emul")
    host = os.path.join(ROOT, "hal_b200", "csrc", "host")
    srcs = [os.path.join(host, f) for f in ("halWiggleLiftoverMain.cpp", "wiggle_liftover.cpp")]
    deps = srcs + [os.path.join(host, "wiggle_liftover.hpp"), emul_lib]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", out] + srcs +
                              ["-L" + os.path.dirname(emul_lib), "-lhalgpu_emul", "-Wl,-rpath,$ORIGIN", "-pthread"])
    return out


@pytest.fixture(scope="session")
def emul_synteny_cli(emul_lib):
    """The halSynteny CLI linked against the emulated library."""
    out = os.path.join(ROOT, "tests", "simt", "halSynteny_emul")
    host = os.path.join(ROOT, "hal_b200", "csrc", "host")
    srcs = [os.path.join(host, f) for f in ("halSyntenyMain.cpp", "synteny.cpp")]
    deps = srcs + [os.path.join(host, "synteny.hpp"), emul_lib]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", out] + srcs +
                              ["-L" + os.path.dirname(emul_lib), "-lhalgpu_emul", "-Wl,-rpath,$ORIGIN", "-pthread"])
    return out


@pytest.fixture(scope="session")
def emul_blockviz_cli(emul_lib):
    """tests/cpp/blockviz_cli.cpp against this repo's blockViz implementation linked with the emulated library."""
    d = os.path.join(ROOT, "tests", "simt")
    lib, out = os.path.join(d, "libhalBlockVizGpu_emul.so"), os.path.join(d, "blockVizCli_emul")
    host = os.path.join(ROOT, "hal_b200", "csrc", "host")
    srcs = [os.path.join(host, "blockviz.cpp"), os.path.join(host, "maf_export.cpp")]
    drv = os.path.join(ROOT, "tests", "cpp", "blockviz_cli.cpp")
    deps = srcs + [drv, os.path.join(host, "maf_export.hpp"), os.path.join(ROOT, "include", "halgpu_blockviz.h"), emul_lib]
    if not os.path.exists(out) or any(os.path.getmtime(x) > os.path.getmtime(out) for x in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", lib] + srcs + ["-L" + d, "-lhalgpu_emul", "-Wl,-rpath,$ORIGIN", "-pthread"])
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-DHALGPU_BLOCKVIZ_HEADER", "-I" + os.path.join(ROOT, "include"), "-o", out, drv, "-L" + d, "-lhalBlockVizGpu_emul",
                               "-lhalgpu_emul", "-Wl,-rpath,$ORIGIN", "-pthread"])
    return out


@pytest.fixture(scope="session")
def emul_depth_cli(emul_lib):
    out = os.path.join(ROOT, "tests", "simt", "halAlignmentDepth_emul")
    src = os.path.join(ROOT, "hal_b200", "csrc", "host", "halAlignmentDepthMain.cpp")
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in (src, emul_lib)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", out, src, "-L" + os.path.dirname(emul_lib), "-lhalgpu_emul",
                               "-Wl,-rpath,$ORIGIN", "-pthread"])
    return out


@pytest.fixture(scope="session")
def emul_maf_cli(emul_lib):
    out = os.path.join(ROOT, "tests", "simt", "hal2maf_emul")
    host = os.path.join(ROOT, "hal_b200", "csrc", "host")
    srcs = [os.path.join(host, f) for f in ("hal2mafMain.cpp", "maf_export.cpp", "bed.cpp")]
    deps = srcs + [os.path.join(host, "maf_export.hpp"), os.path.join(host, "bed.hpp"), emul_lib]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", out] + srcs + ["-L" + os.path.dirname(emul_lib), "-lhalgpu_emul",
                               "-Wl,-rpath,$ORIGIN", "-pthread"])
    return out


def ref_bin(name):
    p = os.path.join(ROOT, "oracle", "_ref", name)
    return p if os.path.exists(p) else None
