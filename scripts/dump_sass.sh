#!/bin/bash
# Dumps the SASS of the dominant kernel (the one bench.py times: k_tile_ws<float, 704, 4>) from the built
# library, with the counts of the mnemonics that identify a hand-written Blackwell kernel:
#   UBLKCP            cp.async.bulk (TMA bulk copy, global -> shared)
#   SYNCS             mbarrier arrive / expect_tx / try_wait
#   FFMA2/FMUL2/FADD2 packed f32x2 arithmetic (sm_100)
#   LDS/STS           shared-memory gathers and the dv read-modify-write
# Usage: scripts/dump_sass.sh [function-substring] > profiles/rNN_tile_tma_sass.txt
set -euo pipefail
SO="$(dirname "$0")/../odesat_b200/csrc/libodesat_b200.so"
FN="${1:-_ZN6odesat9k_tile_wsIfLi704ELi4EEEvNS_8TileArgsIT_EENS_8TileWorkE}"
TMP="$(mktemp)"
cuobjdump -sass -fun "$FN" "$SO" > "$TMP"
echo "# cuobjdump -sass -fun $FN libodesat_b200.so   ($(date -u +%Y-%m-%dT%H:%MZ), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
echo "# arch: $(grep -m1 -o 'sm_[0-9a-z]*' "$TMP" || true)"
echo "# instruction counts:"
for m in UBLKCP SYNCS FFMA2 FMUL2 FADD2 FFMA FMUL FADD FMNMX LDS STS LDG STG LDGSTS BAR MEMBAR FENCE ATOM RED; do
  printf "#   %-8s %s\n" "$m" "$(grep -cE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T]+\s+)?$m(\.|\s|;)" "$TMP" || true)"
done
echo "# total instructions: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' "$TMP")"
cat "$TMP"
rm -f "$TMP"
