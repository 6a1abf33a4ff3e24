#!/bin/sh
# Pins the oracle against the real bendy2d crate:  oracle/ref_harness/run.sh [/path/to/bendy2d]   (needs cargo + network
# or a vendored nalgebra 0.32.2).  Builds the harness against the UNMODIFIED reference, runs the three golden scenes and
# compares the final states with tests/golden/*.npz bit for bit.  Never run in the build image (no Rust toolchain).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${1:-/root/reference}
WORK=$(mktemp -d)
cp -r "$HERE/src" "$WORK/"
sed "s#REFERENCE_PATH#$REF#" "$HERE/Cargo.toml" > "$WORK/Cargo.toml"
python "$HERE/export_inputs.py" "$WORK/inputs"
(cd "$WORK" && cargo build --release)
mkdir -p "$WORK/outputs"
for s in c1_reference_order circle_pile polygon_heap; do
  "$WORK/target/release/bendy2d-ref-harness" "$WORK/inputs/$s.txt" > "$WORK/outputs/$s.txt"
done
python "$HERE/compare.py" "$WORK/outputs"
