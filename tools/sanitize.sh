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
g++ $F -DHALGPU_SIMT_EMUL -I$R/tests/simt -I$C -fPIC -shared -pthread -x c++ $C/capi.cu -x c++ $C/engine.cu -x c++ $C/multi.cu -x c++ $C/halmmap.cpp -o libhalgpu_emul.so
L="-L. -lhalgpu_emul -Wl,-rpath,\$ORIGIN -pthread"
g++ $F -o halLiftover_emul $H/halLiftoverMain.cpp $H/gpu_liftover.cpp $H/bed.cpp $H/bed_fast.cpp $L
g++ $F -o halWiggleLiftover_emul $H/halWiggleLiftoverMain.cpp $H/wiggle_liftover.cpp $L
g++ $F -o halSynteny_emul $H/halSyntenyMain.cpp $H/synteny.cpp $L
g++ $F -fPIC -shared -o libhalBlockVizGpu_emul.so $H/blockviz.cpp $H/maf_export.cpp $L
g++ $F -DHALGPU_BLOCKVIZ_HEADER -I$R/include -o blockVizCli_emul $R/tests/cpp/blockviz_cli.cpp -L. -lhalBlockVizGpu_emul $L
g++ $F -o halAlignmentDepth_emul $H/halAlignmentDepthMain.cpp $L
g++ $F -o hal2maf_emul $H/hal2mafMain.cpp $H/maf_export.cpp $H/bed.cpp $L
export ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 HALGPU_CLEAN_EXIT=1 HALGPU_TEXT_THREADS=4 HALGPU_WIG_GRAIN=50
cd "$R"
for hal in tests/golden/refBedLiftoverTest.hal tests/golden/varlen8.hal; do
    python tools/wig_cli_vs_ref.py $A/halWiggleLiftover_emul $hal 1 | tail -n 1
    python tools/synteny_cli_vs_ref.py $A/halSynteny_emul $hal | tail -n 1
    python tools/blockviz_vs_ref.py $A/blockVizCli_emul $hal 40 3 | tail -n 1
    python tools/maf_targets_vs_ref.py $A/hal2maf_emul $hal 1 | tail -n 1
    python tools/cli_sanitized_vs_plain.py $A $hal | tail -n 1
done

# ThreadSanitizer over the multi-threaded text layers (BED tokeniser/printer against the stub ABI library, MAF row pool and
# wiggle scanner/writer against the plain emulated library): no report, output equal to the serial path.
T=$A/tsan; mkdir -p "$T"; cd "$T"
FT="-std=c++17 -O1 -g -fsanitize=thread"
cp $R/tests/simt/libhalgpu_emul.so .
g++ -std=c++17 -O2 -fPIC -shared -o libhalgpu_stub.so $R/tests/cpp/halgpu_stub.cpp
g++ $FT -pthread -o halLiftover_stub $H/halLiftoverMain.cpp $H/gpu_liftover.cpp $H/bed.cpp $H/bed_fast.cpp -L. -lhalgpu_stub '-Wl,-rpath,$ORIGIN'
g++ $FT -o hal2maf_emul $H/hal2mafMain.cpp $H/maf_export.cpp $H/bed.cpp $L
g++ $FT -o halWiggleLiftover_emul $H/halWiggleLiftoverMain.cpp $H/wiggle_liftover.cpp $L
export TSAN_OPTIONS="halt_on_error=1 report_signal_unsafe=0"
python - <<'PY'
import random
rng = random.Random(1)
with open("in.bed", "w") as f:
    for i in range(200000):
        a = rng.randrange(0, 900000)
        f.write(f"{rng.choice(['chrA', 'chrB'])}\t{a}\t{a + rng.randrange(1, 500)}\tn{i}\t{rng.randrange(1000)}\t{rng.choice('+-')}\n")
PY
HALGPU_TEXT_THREADS=6 HALGPU_BLOCK_BYTES=1500000 ./halLiftover_stub x S in.bed T out.bed
HALGPU_TEXT_THREADS=0 ./halLiftover_stub x S in.bed T out0.bed
cmp out.bed out0.bed && echo "tsan BED: clean"
G=$R/tests/golden
HALGPU_TEXT_THREADS=5 HALGPU_MAF_QUEUE_BYTES=20000 ./hal2maf_emul $G/varlen8.hal o.maf --refGenome L0
$R/tests/simt/hal2maf_emul $G/varlen8.hal o0.maf --refGenome L0
cmp o.maf o0.maf && echo "tsan MAF: clean"
