#!/bin/bash
# Build the reference's own Java arithmetic headless (needs a JDK >= 8 on PATH: javac, java; python3).
# Inputs : the UNMODIFIED tabs of <reference>/clientLilypad (default /root/reference/clientLilypad), read where they lie.
# Outputs: oracle/_ref/ only (git-ignored): LilypadSketch.java (generated), *.class, and the golden files.
# The build container of this repository has no JDK, so this recipe has not been run there (see README.md).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference/clientLilypad}"
OUT="$HERE/../_ref"
command -v javac >/dev/null || { echo "build_ref.sh: no javac on PATH (the reference cannot be built here)"; exit 3; }
mkdir -p "$OUT"
TABS=(Window OrthoNormal Body BodyUnion Field VectorField PoissonMatrix MG BDIM SaveScalar AFCCylinder)
(cd "$REF" && python3 "$HERE/pde2java.py" "$OUT/LilypadSketch.java" "${TABS[@]/%/.pde}")
# strictfp is the default from JDK 17 on; older JDKs on x86-64 use SSE arithmetic, which is IEEE binary32 for float as well
javac -nowarn -d "$OUT" "$HERE/PAppletShim.java" "$OUT/LilypadSketch.java" "$HERE/HeadlessLilypad.java"
java -cp "$OUT" HeadlessLilypad "$REF" "$OUT/golden"
python3 "$HERE/compare_ref.py" "$OUT/golden"
