#!/usr/bin/env bash
# Pin this repository's oracle (and CUDA library) against the REAL reference implementation.
#
# Needs what the build image lacks: cargo/rustc, network access for the un-vendored crates (ICICLE v3.8.0, arkworks 0.5,
# tiny-keccak) and the ICICLE CPU backend libraries.  Run it once on such a machine:
#
#   scripts/pin/pin_against_reference.sh /path/to/Tokamak-zk-EVM
#
# It (1) drops scripts/pin/pin_harness.rs into packages/backend/libs/tests/, (2) runs it with the reference's own
# library code and prints NTT / root-of-unity / MSM known answers, (3) stores them as tests/golden/reference_pins.json,
# (4) optionally produces a fixed-tau CRS and a proof with the reference's binaries (trusted-setup --fixed-tau, preprocess,
# prove; setup/trusted-setup/src/main.rs:71-78, prove/optimization/tests/timing.rs:98-233) into tests/golden/reference_flow/
# so that the byte-level proof comparison can be made as well.  tests/test_reference_pins.py consumes whatever is present;
# until this has been run the oracle's header and DESIGN.md say "parity unpinned".
set -euo pipefail
REF="${1:?path to a checkout of tokamak-network/Tokamak-zk-EVM}"
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
BACKEND="$REF/packages/backend"
command -v cargo >/dev/null || { echo "cargo not found: this script must run where the reference builds" >&2; exit 2; }
cp "$HERE/pin_harness.rs" "$BACKEND/libs/tests/pin_harness.rs"
trap 'rm -f "$BACKEND/libs/tests/pin_harness.rs"' EXIT
OUT="$(cd "$BACKEND" && cargo test --release -p libs --test pin_harness -- --nocapture --test-threads=1)"
echo "$OUT" | sed -n '/-----BEGIN REFERENCE PINS-----/,/-----END REFERENCE PINS-----/p' | sed '1d;$d' > "$REPO/tests/golden/reference_pins.json"
python3 -c "import json,sys; d=json.load(open('$REPO/tests/golden/reference_pins.json')); print('pinned:', sorted(d))"
if [ "${PIN_FLOW:-0}" = "1" ]; then
  # full flow on the reference's own binaries with the fixed trapdoor; the synthesizer outputs must be supplied by the caller
  # (SYN_DIR: placementVariables.json, permutation.json, instance.json) because the synthesizer is a Node package
  : "${SYN_DIR:?set SYN_DIR to a synthesizer output directory}"
  FLOW="$REPO/tests/golden/reference_flow"; mkdir -p "$FLOW"
  QAP="$REF/packages/frontend/qap-compiler/subcircuits/library"
  (cd "$BACKEND" && cargo run --release -p trusted-setup -- "$QAP" "$FLOW" --fixed-tau)
  (cd "$BACKEND" && cargo run --release -p preprocess -- "$QAP" "$SYN_DIR" "$FLOW" "$FLOW")
  (cd "$BACKEND" && cargo run --release -p prove -- "$QAP" "$SYN_DIR" "$FLOW" "$FLOW")
  cp "$SYN_DIR"/{placementVariables.json,permutation.json,instance.json} "$FLOW"/
  echo "reference flow artefacts in $FLOW (proof.json uses the reference's random mixer: compare commitments that do not depend on it, and verify)"
fi
echo "now run: python -m pytest tests/test_reference_pins.py -q"
