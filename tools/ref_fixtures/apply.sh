#!/bin/bash
# Generates tests/golden/ref/*.json from a checkout of the reference (see README.md).  Needs cargo; NOT runnable in the build container.
set -euo pipefail
ref=${1:?usage: apply.sh /path/to/bevy_terrain-checkout}
here=$(cd "$(dirname "$0")" && pwd)
out=$(cd "$here/../.." && pwd)/tests/golden/ref
work=$(mktemp -d)
cp -r "$ref"/. "$work"/
cp "$here/bt_ref_fixtures.rs" "$work/src/bt_ref_fixtures.rs"
printf '\n#[cfg(test)]\nmod bt_ref_fixtures;\n' >> "$work/src/lib.rs"
# compute_origin / compute_tile_distance are private to terrain_data::tile_tree: their dumper has to live inside that module
cat "$here/bt_ref_fixtures_tile_tree.rs" >> "$work/src/terrain_data/tile_tree.rs"
mkdir -p "$out"
(cd "$work" && BT_REF_FIXTURES_DIR="$out" cargo test --lib bt_ref_fixtures -- --nocapture --test-threads 1)
echo "fixtures written to $out:"; ls -la "$out"
