#!/usr/bin/env bash
# Build the UNMODIFIED reference (read-only at /root/reference) into oracle/_ref/.
#
# TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path; it is the
# checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
#
# The reference's own CMake needs find_package(BLAS), which fails in this image (no system
# BLAS/LAPACK), so the sources are compiled where they lie with gcc against the OpenBLAS that
# ships inside the venv (recipe: SURVEY.md section 8(c)).  No reference source is copied:
# only object code lands in oracle/_ref/ (git-ignored, but it travels to the GPU box).
#
#   oracle/_ref/libhdsdp_ref.so   all reference objects + oracle/ref_driver.c (ctypes-callable)
#   oracle/_ref/sdpasolve         the reference CLI (tests/sdpasolve.c)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${HDSDP_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
PY="${PYTHON:-python}"
if [ ! -d "$REF/interface" ]; then
    echo "build_ref: $REF not present; keeping prebuilt oracle/_ref" >&2
    exit 0
fi
SITE="$($PY - <<'PY'
import sysconfig; print(sysconfig.get_paths()["purelib"])
PY
)"
OBDIR="$SITE/opencv_python_headless.libs"
OBLIB="$(ls "$OBDIR"/libopenblasp-*.so 2>/dev/null | head -1 || true)"
if [ -z "$OBLIB" ]; then echo "build_ref: no OpenBLAS found under $OBDIR" >&2; exit 1; fi
mkdir -p "$OUT/obj"
CFLAGS="-O2 -std=gnu99 -DHEADERPATH -DUNDERBLAS -I$REF -fPIC -w"
for f in "$REF"/interface/*.c "$REF"/linalg/*.c "$REF"/external/*.c; do
    o="$OUT/obj/$(basename "${f%.c}").o"
    if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then gcc $CFLAGS -c "$f" -o "$o"; fi
done
gcc $CFLAGS -c "$HERE/ref_driver.c" -o "$OUT/obj/ref_driver.o"
gcc -shared -o "$OUT/libhdsdp_ref.so" "$OUT"/obj/*.o "$OBLIB" -Wl,--disable-new-dtags,-rpath,"$OBDIR" -lm
# tests/sdpasolve.c #includes tests/test_file_io.c itself
gcc $CFLAGS "$REF/tests/sdpasolve.c" $(ls "$OUT"/obj/*.o | grep -v ref_driver.o) \
    "$OBLIB" -Wl,--disable-new-dtags,-rpath,"$OBDIR" -lm -o "$OUT/sdpasolve"
echo "$OBLIB" > "$OUT/BLAS_USED.txt"
echo "build_ref: built $OUT/libhdsdp_ref.so and $OUT/sdpasolve against $(basename "$OBLIB")"
