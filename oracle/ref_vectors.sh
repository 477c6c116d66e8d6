#!/bin/bash
# Golden vectors from the reference binary built by oracle/build_ref.sh (never run in this container: see there).
# Writes the fixture world of tests/sv_world.py to a scratch directory, runs the reference's own pipeline on it
# (run_svdss:136-178: index -> search -> call; the BAM of the fixture is already smoothed-shaped, and `samtools index`
# must be on PATH for `call`), and keeps what the tests compare: the .sfs text, the cluster file, the VCF body, and a
# ropebwt3-written .fmd of the fixture reference (pins svdss_b200/host/rld.hpp).
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(dirname "$HERE")
BIN=$HERE/_ref/SVDSS
[ -x "$BIN" ] || { echo "run oracle/build_ref.sh first"; exit 1; }
W=$(mktemp -d)
( cd "$ROOT" && PYTHONPATH=$ROOT:$ROOT/tests python -c "
import sys
from sv_world import make_world
w = make_world('$W')
print(w['fa'], w['bam'])" )
"$BIN" index -t 4 -d "$W/ref.fa" -o "$W/ref.fmd"
samtools index "$W/sample.bam"
"$BIN" search --threads 4 --index "$W/ref.fmd" --bam "$W/sample.bam" --noputative > "$W/ref.sfs"
"$BIN" call --threads 4 --reference "$W/ref.fa" --bam "$W/sample.bam" --sfs "$W/ref.sfs" --clusters "$W/ref.clusters" > "$W/ref.vcf"
mkdir -p "$ROOT/tests/golden"
cp "$W/ref.fmd" "$ROOT/tests/golden/ref_world.fmd"
cp "$W/ref.sfs" "$ROOT/tests/golden/ref_world.sfs"
cp "$W/ref.clusters" "$ROOT/tests/golden/ref_world.clusters"
grep -v '^#' "$W/ref.vcf" > "$ROOT/tests/golden/ref_world.vcf" || true
echo "wrote tests/golden/ref_world.{fmd,sfs,clusters,vcf}"
