#!/bin/sh
# Builds tests/cpp/dropin_program.cpp against the UNMODIFIED reference headers (+ the oracle's Eigen stand-in)
# and records its output as the expected output of the same program built against include/ + the CUDA library.
# Run in the build container (needs /root/reference).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
REF=${REF:-/root/reference}
OUT=$(mktemp -d)
g++ -std=c++17 -O2 -ffp-contract=off -I"$ROOT/oracle/shim" -I"$REF/bonxai_core/include" -I"$REF/bonxai_map/include" \
    "$ROOT/tests/cpp/dropin_program.cpp" "$REF/bonxai_map/src/probabilistic_map.cpp" -o "$OUT/dropin_ref"
"$OUT/dropin_ref" > "$HERE/dropin_expected.txt"
cat "$HERE/dropin_expected.txt"
