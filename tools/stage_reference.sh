#!/bin/bash
# Stage the UNMODIFIED reference into the git-ignored baseline/_ref (it is NOT gpurun-ignored, so it travels to the GPU
# box): `pip install --no-index --no-deps --target baseline/_ref` of a copy of /root/reference (the tree is read-only and
# setuptools writes build/ + egg-info next to setup.py), plus its config/ directory.  Nothing of it enters git history.
# Run in the build container (the only place /root/reference exists); the GPU-side tests / bench legs that need it skip or
# say "unavailable" when it is absent.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=${1:-/root/reference}
[ -d "$SRC/generalframework" ] || { echo "no reference tree at $SRC"; exit 1; }
TMP=$(mktemp -d)
mkdir -p "$TMP/ref"
cp -r "$SRC/generalframework" "$SRC/setup.py" "$SRC/README.md" "$TMP/ref/"
rm -rf "$ROOT/baseline/_ref"; mkdir -p "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$TMP/ref" 2>&1 | tail -2
cp -r "$SRC/config" "$ROOT/baseline/_ref/config"
rm -rf "$TMP"
( cd "$ROOT/baseline/_ref" && find generalframework -name '*.py' | sort | xargs sha256sum | sha256sum | cut -c1-16 > .tree_sha16 )
( cd "$SRC" && find generalframework -name '*.py' | sort | xargs sha256sum | sha256sum | cut -c1-16 ) > "$ROOT/baseline/_ref/.source_sha16"
echo "staged: $(du -sh "$ROOT/baseline/_ref" | cut -f1), tree sha16 $(cat "$ROOT/baseline/_ref/.tree_sha16") (source $(cat "$ROOT/baseline/_ref/.source_sha16"))"
