#!/usr/bin/env bash
# Usage: tools/compare_with_jar.sh /path/to/sagephy-1.0.0.jar <seed>
# Runs the reference jar and the native CLI (MT replay) on the README examples and diffs every output file.
# Needs a JVM (>= 19 for shortest-digit Double.toString) — not available in the build image.
set -euo pipefail
JAR=$1; SEED=${2:-1}; HERE=$(cd "$(dirname "$0")/.." && pwd); W=$(mktemp -d)
JAVA="java -XX:+UnlockDiagnosticVMOptions -XX:-UseLibmIntrinsic -jar $JAR"
run() { app=$1; shift; (cd "$W" && mkdir -p ref gpu && (cd ref && $JAVA "$app" "$@") && (cd gpu && "$HERE/sagephy_b200/sagephy" "$app" -rng mt-replay "$@")); }
run HostTreeGen -s "$SEED" -min 100 -max 100 1.0 5.0 0.05 species
cp "$W/ref/species.pruned.tree" "$W/species.pruned.tree"
run GuestTreeGen -s "$SEED" -rt 0.6 -db exponential -dbr 1.5 -gb -gbc 2.1 "$W/species.pruned.tree" 0.2 0.1 0.3 gene
run BranchRelaxer -s "$SEED" -x -innms -o gene.relaxed.tree "$W/ref/gene.pruned.tree" IIDLogNormal 0.0 0.25
diff -r "$W/ref" "$W/gpu" && echo "IDENTICAL"
