#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (Rouslan/NTracer, /root/reference)
# into oracle/_ref/ntracer so that tests/golden/make_fixtures.py can generate golden vectors and
# bench.py's cpu_baseline / --impl reference leg can time the reference's own CPU renderer.
# Nothing in ntracer_b200/ imports or links what this produces.
#
# The reference needs three workarounds to compile on this image:
#   1. a PKG-INFO with a version (support/version.py falls back to 'unversioned' without git);
#   2. -march=x86-64-v2 instead of -march=native (its AVX/AVX2/AVX-512 paths do not compile:
#      v_array.hpp:518 vs simd.hpp.in:756-770)  => SSE4.2 flavour, BATCH_SIZE = 4;
#   3. a force-included shim supplying <functional> and _PyObject_GC_Malloc (gone in CPython 3.12).
# The build writes into its source tree, so it runs on a scratch copy; only the built package
# (extension modules + the package's own .py files) is copied to oracle/_ref/.  oracle/_ref/ is
# git-ignored: no reference source is committed to this repository.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${NTRACER_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
DIMS="${NTRACER_REF_DIMS:-3,4,5,6}"
if [ ! -d "$REF/src" ]; then
    echo "build_ref.sh: reference tree not found at $REF (expected on the build container only)" >&2
    exit 3
fi
SCRATCH="$(mktemp -d /tmp/ntref_build.XXXXXX)"
trap 'rm -rf "$SCRATCH"' EXIT
cp -r "$REF" "$SCRATCH/src"
cd "$SCRATCH/src"
printf 'Metadata-Version: 1.0\nName: ntracer\nVersion: 0.0.0\n' > PKG-INFO
cat > "$SCRATCH/shim.h" <<'SHIM'
#ifdef __cplusplus
#include <functional>
#include <Python.h>
static inline void* _PyObject_GC_Malloc(size_t s){ char*p=(char*)PyObject_Calloc(1,s+16); return p? p+16:nullptr; }
#endif
SHIM
python3 setup.py build -j "$(nproc)" --optimize-dimensions="$DIMS" \
    --cpp-neg-opts=-march=native \
    --cpp-opts="-march=x86-64-v2 -include $SCRATCH/shim.h" > "$SCRATCH/build.log" 2>&1 \
    || { tail -40 "$SCRATCH/build.log" >&2; exit 1; }
BUILT="$(ls -d build/lib.*/ntracer)"
rm -rf "$OUT"
mkdir -p "$OUT"
cp -r "$BUILT" "$OUT/ntracer"
# the two demo scripts are scene generators for the benchmark configs (polytope tessellation)
mkdir -p "$OUT/scripts"
cp scripts/polytope.py scripts/hypercube.py "$OUT/scripts/"
PYTHONPATH="$OUT" python3 -m ntracer.tests.test 2>&1 | tail -3
echo "reference built into $OUT"
