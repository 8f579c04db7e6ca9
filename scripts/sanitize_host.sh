#!/bin/bash
# AddressSanitizer + UBSan run of the host-side code (builder, flatten incl. tile table, entry nodes,
# exact replay and fast-path logic of gsi_eval_core.h) on degenerate and random inputs.
set -e
C=$(dirname "$0")/../gsl_scattered_interpolation_b200/csrc
OUT=${TMPDIR:-/tmp}/gsi_sanitize_host
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -std=c++17 -I"$C" \
    -o "$OUT" "$(dirname "$0")/sanitize_host.cpp" "$C/gsi_builder.cpp" "$C/gsi_flatten.cpp"
ASAN_OPTIONS=detect_leaks=1 "$OUT"
