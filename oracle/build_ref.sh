#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds the UNMODIFIED reference (nickmvincent/Surprise, Cython) into
# oracle/_ref/ so that (a) the C restatement in oracle/ can be pinned against it and (b)
# `bench.py --impl reference` can time the reference's own code path on the GPU box's host cores.
#
# /root/reference is read-only, so the build runs on a scratch copy under /tmp. Three mechanical
# toolchain-compat edits are applied to that copy (numpy>=1.24 removed np.int, Cython 3 rejects
# np.int_t, setuptools rejects the version string 'latest', the package is not pip-registered);
# none touches arithmetic.  Nothing from the reference is committed: oracle/_ref/ is git-ignored.
set -euo pipefail
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/surprise" ]; then
  echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt $OUT" >&2
  exit 0
fi
TMP=$(mktemp -d /tmp/surprise_ref_build.XXXXXX)
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF"/. "$TMP"/
chmod -R u+w "$TMP"
cd "$TMP"
# 1. np.int_t -> np.int64_t, np.int) -> np.int64)  (similarities / slope_one / co_clustering)
sed -i -e 's/np\.int_t/np.int64_t/g' -e 's/np\.int)/np.int64)/g' -e 's/np\.int,/np.int64,/g' \
    surprise/similarities.pyx surprise/prediction_algorithms/slope_one.pyx \
    surprise/prediction_algorithms/co_clustering.pyx
# 2. version strings
sed -i "s/__version__ = 'latest'/__version__ = '1.0.5'/" setup.py
sed -i "s/^__version__ = get_distribution('scikit-surprise').version/__version__ = '1.0.5'/" surprise/__init__.py
python setup.py -q build_ext --inplace >"$TMP/build.log" 2>&1 || { tail -50 "$TMP/build.log"; exit 1; }
rm -rf "$OUT"
mkdir -p "$OUT"
cp -r surprise "$OUT"/surprise
find "$OUT" \( -name '*.c' -o -name '*.pyx' \) -delete
find "$OUT" -depth -name '__pycache__' -exec rm -rf {} +
echo "build_ref: reference built into $OUT"
