#!/bin/bash
# Sanitizer tier (the reference runs its suites under ASan in CI, .travis.yml:23-31): builds the kernel/engine sources for the
# warp emulator and every host CLI / library with -fsanitize=address,undefined into a scratch directory and drives them
# through the reference-comparison tools.  Slow (one OS thread per CUDA thread under ASan): tens of minutes on 8 cores.
# usage: tools/sanitize.sh [scratch dir]      (needs oracle/_ref, i.e. /root/reference at build time)
set -euo pipefail
R=$(cd "$(dirname "$0")/.." && pwd)
A=${1:-/tmp/halgpu_asan}
C=$R/hal_b200/csrc; H=$C/host
F="-std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer"
mkdir -p "$A"; cd "$A"
g++ $F -DHALGPU_SIMT_EMUL -I$R/tests/simt -I$C -fPIC -shared -pthread -x c++ $C/capi.cu -x c++ $C/engine.cu -x c++ $C/halmmap.cpp -o libhalgpu_emul.so
L="-L. -lhalgpu_emul -Wl,-rpath,\$ORIGIN -pthread"
g++ $F -o halLiftover_emul $H/halLiftoverMain.cpp $H/gpu_liftover.cpp $H/bed.cpp $H/bed_fast.cpp $L
g++ $F -o halWiggleLiftover_emul $H/halWiggleLiftoverMain.cpp $H/wiggle_liftover.cpp $L
g++ $F -o halSynteny_emul $H/halSyntenyMain.cpp $H/synteny.cpp $L
g++ $F -fPIC -shared -o libhalBlockVizGpu_emul.so $H/blockviz.cpp $H/maf_export.cpp $L
g++ $F -DHALGPU_BLOCKVIZ_HEADER -I$R/include -o blockVizCli_emul $R/tests/cpp/blockviz_cli.cpp -L. -lhalBlockVizGpu_emul $L
g++ $F -o halAlignmentDepth_emul $H/halAlignmentDepthMain.cpp $L
g++ $F -o hal2maf_emul $H/hal2mafMain.cpp $H/maf_export.cpp $H/bed.cpp $L
export ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 HALGPU_TEXT_THREADS=4 HALGPU_WIG_GRAIN=50
cd "$R"
for hal in tests/golden/refBedLiftoverTest.hal tests/golden/varlen8.hal; do
    python tools/wig_cli_vs_ref.py $A/halWiggleLiftover_emul $hal 1 | tail -n 1
    python tools/synteny_cli_vs_ref.py $A/halSynteny_emul $hal | tail -n 1
    python tools/blockviz_vs_ref.py $A/blockVizCli_emul $hal 40 3 | tail -n 1
    python tools/maf_targets_vs_ref.py $A/hal2maf_emul $hal 1 | tail -n 1
    python tools/cli_sanitized_vs_plain.py $A $hal | tail -n 1
done
