#!/bin/bash
# Mid-scale parity away from the unit fixtures: the bench's C2 tree at 16 x 3.2 Mbp with halSynth --branch 0.05,
# this repo's CLIs (emulated kernels from tests/simt, or pass a directory of CUDA builds) against the reference binaries in oracle/_ref.
# Every reference call runs under `timeout`: the reference's halSynteny needs 4-10 minutes per genome pair at this size (its dag_merge
# re-weighs the whole DAG per chain), which is most of this script's ~15 minutes.
# usage: tools/midscale_check.sh [dir with the CLIs to test (default tests/simt, names *_emul)] [suffix (default _emul)]
set -u
R=$(cd "$(dirname "$0")/.." && pwd)
E=${1:-$R/tests/simt}; SFX=${2-_emul}
REF=$R/oracle/_ref
W=$(mktemp -d)
H=$W/c2div.hal
NEWICK=$(cd "$R" && python -c "import bench; print(bench.NEWICK)")
$R/hal_b200/bin/halSynth --newick "$NEWICK" --segs 100000 --segLen 32 --branch 0.05 --seed 7 $H
bad=0
same() { if cmp -s "$1" "$2"; then echo "SAME  $3"; else echo "DIFF  $3"; bad=$((bad + 1)); fi; }
while read -r args; do
    timeout 300 $REF/hal2maf $H $W/a.maf $args; timeout 600 $E/hal2maf$SFX $H $W/b.maf $args; same $W/a.maf $W/b.maf "hal2maf $args"
done <<'LIST'
--refGenome L0 --refSequence L0_seq --start 1000000 --length 60000
--refGenome R --refSequence R_seq --start 500000 --length 60000 --noDupes
--refGenome A1 --refSequence A1_seq --start 2000000 --length 50000 --unique
--refGenome L7 --refSequence L7_seq --start 0 --length 40000 --onlyOrthologs --noAncestors
LIST
while read -r args; do
    timeout 300 $REF/halAlignmentDepth $H $args > $W/a.wig 2>/dev/null; timeout 600 $E/halAlignmentDepth$SFX $H $args > $W/b.wig 2>/dev/null
    same $W/a.wig $W/b.wig "halAlignmentDepth $args"
done <<'LIST'
L0 --start 1000000 --length 80000
R --countDupes --start 5000 --length 80000
A2 --noAncestors --start 2000000 --length 50000 --step 3
LIST
while read -r q t args; do
    if timeout 400 $REF/halSynteny --queryGenome $q --targetGenome $t $args $H $W/a.psl 2>/dev/null; then
        timeout 600 $E/halSynteny$SFX --queryGenome $q --targetGenome $t $args $H $W/b.psl 2>/dev/null; same $W/a.psl $W/b.psl "halSynteny $q $t $args"
    else
        echo "SKIP  halSynteny $q $t $args (reference exceeded 400 s)"
    fi
done <<'LIST'
L4 L5 --minBlockSize 100 --maxAnchorDistance 500
R L5 --minBlockSize 50 --maxAnchorDistance 2000
LIST
python - "$W" <<'PY'
import sys
import numpy as np
rng = np.random.default_rng(3)
with open(sys.argv[1] + "/in.wig", "w") as f:
    f.write("fixedStep chrom=L7_seq start=1000001 step=1\n" + "".join(f"{x:.3f}\n" for x in rng.random(60000) * 50))
    f.write("variableStep chrom=L7_seq span=3\n" + "".join(f"{2000000 + 7 * i}\t{(i % 13) - 4}\n" for i in range(8000)))
PY
for nd in "" "--noDupes"; do
    timeout 300 $REF/halWiggleLiftover $nd $H L7 $W/in.wig L0 $W/a.wig; timeout 600 $E/halWiggleLiftover$SFX $nd $H L7 $W/in.wig L0 $W/b.wig
    same $W/a.wig $W/b.wig "halWiggleLiftover L7 L0 $nd"
done
echo "$bad mismatches"; rm -rf "$W"; exit $bad
