#!/usr/bin/env bash
# Usage: tools/time_reference_jar.sh /path/to/sagephy-1.0.0.jar [n_trees] [first_seed]
#
# Times the reference jar on the BASELINE configs on a box that has a JVM (none exists in the build image or on the GPU boxes,
# BASELINE.md section 3.2) and diff-checks every file it writes against the CPU oracle for the same seeds, which is the one
# route to pinning oracle == jar.
#   * "per launch": one `java -jar` per tree, as the reference ships (apps/SaGePhyStarter.java:36-47 scans the classpath on every
#     start) — the cost a user of the reference pays per tree;
#   * "amortised" : wall time of the loop minus n x the start-up time of a no-op launch (`HostTreeGen -h`).
# Prints trees/s for both and the number of files that differ from the oracle (0 = pinned for these seeds).
# JDK >= 19 is needed for shortest-digit Double.toString; -XX:-UseLibmIntrinsic keeps Math.log/exp/pow on fdlibm.
set -euo pipefail
JAR=${1:?path to sagephy-1.0.0.jar}; N=${2:-20}; SEED=${3:-1}
HERE=$(cd "$(dirname "$0")/.." && pwd); W=$(mktemp -d)
ORACLE="$HERE/oracle/sagephy_oracle"; [ -x "$ORACLE" ] || make -C "$HERE/oracle"
JAVA=(java -XX:+UnlockDiagnosticVMOptions -XX:-UseLibmIntrinsic -jar "$JAR")
now() { date +%s.%N; }

t0=$(now); for _ in $(seq 5); do "${JAVA[@]}" HostTreeGen -h > /dev/null 2>&1 || true; done; t1=$(now)
START=$(echo "($t1 - $t0) / 5" | bc -l)
printf 'JVM start-up (no-op launch): %.3f s\n' "$START"

bench() {   # name, app, args... (the token SEED is replaced by the seed, OUT by the output prefix)
    local name=$1 app=$2; shift 2
    mkdir -p "$W/$name/ref" "$W/$name/ora"
    local t0 t1 s a bad=0
    t0=$(now)
    for i in $(seq 0 $((N - 1))); do
        s=$((SEED + i)); a=("${@//SEED/$s}"); a=("${a[@]//OUT/t$i}")
        (cd "$W/$name/ref" && "${JAVA[@]}" "$app" "${a[@]}" > "t$i.stdout" 2> /dev/null) || true
    done
    t1=$(now)
    for i in $(seq 0 $((N - 1))); do
        s=$((SEED + i)); a=("${@//SEED/$s}"); a=("${a[@]//OUT/t$i}")
        (cd "$W/$name/ora" && "$ORACLE" "$app" "${a[@]}" > "t$i.stdout" 2> /dev/null) || true
    done
    bad=$(diff -rq "$W/$name/ref" "$W/$name/ora" | wc -l)
    local wall; wall=$(echo "$t1 - $t0" | bc -l)
    printf '%-28s per launch %8.2f trees/s   amortised %8.2f trees/s   files differing from the oracle: %d\n' "$name" \
        "$(echo "$N / $wall" | bc -l)" "$(echo "$N / ($wall - $N * $START + 0.000001)" | bc -l)" "$bad"
}

bench C1_host_readme   HostTreeGen  -s SEED -min 100 -max 100 1.0 5.0 0.05 OUT
bench C2_host_p08      HostTreeGen  -s SEED -min 100 -max 100 -p 0.8 -a 10000 1.0 5.0 0.05 OUT
bench C3_guest_none    GuestTreeGen -s SEED -db none "$HERE/tests/golden/species100.pruned.tree" 0.3 0.3 0.3 OUT
bench C3_guest_simple  GuestTreeGen -s SEED "$HERE/tests/golden/species100.pruned.tree" 0.3 0.3 0.3 OUT
TREE=$(cat "$W/C2_host_p08/ref/t0.pruned.tree" 2> /dev/null || cat "$W/C2_host_p08/ora/t0.pruned.tree")
bench C5_relax_lognormal BranchRelaxer -s SEED -o OUT.relaxed "$TREE" IIDLogNormal 0 0.25
bench C5_relax_gamma     BranchRelaxer -s SEED -o OUT.relaxed "$TREE" IIDGamma 2 0.5
echo "work dir: $W"
