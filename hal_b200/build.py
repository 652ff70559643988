"""Build the product artefacts in-tree (they travel to the GPU box with the snapshot):

  hal_b200/libhalgpu.so   C-ABI library, CUDA kernels compiled for sm_100a (nvcc cross-compiles without a GPU)
  hal_b200/bin/halLiftover  the reference CLI's GPU build (host C++ over the C ABI)
  hal_b200/bin/halWiggleLiftover, halAlignmentDepth, hal2maf   likewise
  hal_b200/bin/halSynth   synthetic HAL-MMAP writer (host C++)

Run: python -m hal_b200.build
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhalgpu.so")
BIN = os.path.join(HERE, "bin")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-Wno-deprecated-declarations", "-shared", "-ldl"]
LIB_SOURCES = ["capi.cu", "engine.cu", "multi.cu", "halmmap.cpp"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _all_sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out.append(os.path.join(os.path.dirname(HERE), "include", "halgpu.h"))
    return out


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = _all_sources()
    os.makedirs(BIN, exist_ok=True)
    if force or _newer(LIB, srcs):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
              [os.path.join(CSRC, s) for s in LIB_SOURCES]
        subprocess.check_call(cmd)
    synth = os.path.join(BIN, "halSynth")
    if force or _newer(synth, [os.path.join(CSRC, "host", "halsynth.cpp")]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", synth, os.path.join(CSRC, "host", "halsynth.cpp")])
    host = os.path.join(CSRC, "host")
    cli = os.path.join(BIN, "halLiftover")
    cli_srcs = [os.path.join(host, f) for f in ("halLiftoverMain.cpp", "gpu_liftover.cpp", "bed.cpp", "bed_fast.cpp")]
    if force or _newer(cli, cli_srcs + [os.path.join(host, f) for f in ("gpu_liftover.hpp", "bed.hpp", "bed_fast.hpp")] + [LIB]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-o", cli] + cli_srcs +
                              ["-L" + HERE, "-lhalgpu", "-Wl,-rpath,$ORIGIN/.."])
    wig = os.path.join(BIN, "halWiggleLiftover")
    wig_srcs = [os.path.join(host, f) for f in ("halWiggleLiftoverMain.cpp", "wiggle_liftover.cpp")]
    if force or _newer(wig, wig_srcs + [os.path.join(host, "wiggle_liftover.hpp"), LIB]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-o", wig] + wig_srcs + ["-L" + HERE, "-lhalgpu", "-Wl,-rpath,$ORIGIN/.."])
    syn = os.path.join(BIN, "halSynteny")
    syn_srcs = [os.path.join(host, f) for f in ("halSyntenyMain.cpp", "synteny.cpp")]
    if force or _newer(syn, syn_srcs + [os.path.join(host, "synteny.hpp"), LIB]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", syn] + syn_srcs + ["-L" + HERE, "-lhalgpu", "-Wl,-rpath,$ORIGIN/.."])
    viz = os.path.join(HERE, "libhalBlockVizGpu.so")  # the reference's blockViz C API (include/halgpu_blockviz.h) over libhalgpu
    viz_srcs = [os.path.join(host, "blockviz.cpp"), os.path.join(host, "maf_export.cpp")]
    if force or _newer(viz, viz_srcs + [os.path.join(host, "maf_export.hpp"), os.path.join(os.path.dirname(HERE), "include", "halgpu_blockviz.h"), LIB]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-o", viz] + viz_srcs + ["-L" + HERE, "-lhalgpu", "-Wl,-rpath,$ORIGIN"])
    vizcli = os.path.join(BIN, "blockVizCli")  # text driver of the C API (tests/cpp/blockviz_cli.cpp)
    vizcli_src = os.path.join(os.path.dirname(HERE), "tests", "cpp", "blockviz_cli.cpp")
    if os.path.exists(vizcli_src) and (force or _newer(vizcli, [vizcli_src, viz])):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-DHALGPU_BLOCKVIZ_HEADER", "-I" + os.path.join(os.path.dirname(HERE), "include"), "-o", vizcli, vizcli_src,
                               "-L" + HERE, "-lhalBlockVizGpu", "-lhalgpu", "-Wl,-rpath,$ORIGIN/.."])
    dep = os.path.join(BIN, "halAlignmentDepth")
    dep_src = os.path.join(host, "halAlignmentDepthMain.cpp")
    if force or _newer(dep, [dep_src, LIB]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", dep, dep_src, "-L" + HERE, "-lhalgpu", "-Wl,-rpath,$ORIGIN/.."])
    maf = os.path.join(BIN, "hal2maf")
    maf_srcs = [os.path.join(host, f) for f in ("hal2mafMain.cpp", "maf_export.cpp", "bed.cpp")]
    if force or _newer(maf, maf_srcs + [os.path.join(host, "maf_export.hpp"), LIB]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-o", maf] + maf_srcs + ["-L" + HERE, "-lhalgpu", "-Wl,-rpath,$ORIGIN/.."])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
