#!/usr/bin/env bash
# Patch-free integration build: the UNMODIFIED reference host solver (objects compiled by oracle/build_ref.sh from
# /root/reference) linked against the CUDA hot path, with the three hook files of this directory:
#   hdsdp_schur_cuda.c   replaces interface/hdsdp_schur.c        (its object is simply left out of the link)
#   hdsdp_linsys_cuda.c  wraps HFpLinsysCreate                    (the reference symbol is renamed with objcopy)
#   hdsdp_conic_cuda.c   wraps HConeSetData: the SDP cone vtable points at the device (renamed with objcopy);
#                        the DIMACS eigenvalue call of hdsdp.o (fds_syev) is renamed to the hook's fds_syev_dimacs
# Outputs (git-ignored; they contain reference object code; they travel to the GPU box):
#   integration/_build/libhdsdp_integrated.so   reference + hooks + oracle/ref_driver.c (ctypes entry: refdrv_optimize)
#   integration/_build/sdpasolve_cuda            the reference CLI running on the GPU hot path
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${HDSDP_REFERENCE:-/root/reference}"
OBJ="$ROOT/oracle/_ref/obj"
OUT="$HERE/_build"
if [ ! -d "$REF/interface" ]; then echo "build_integrated: $REF not present; keeping prebuilt integration/_build" >&2; exit 0; fi
[ -f "$OBJ/hdsdp_linsolver.o" ] || bash "$ROOT/oracle/build_ref.sh"
[ -f "$ROOT/hdsdp_b200/libhdsdp_cuda.so" ] || make -C "$ROOT/hdsdp_b200/csrc" -j8
mkdir -p "$OUT"
PY="${PYTHON:-python}"
SITE="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["purelib"])')"
OBDIR="$SITE/opencv_python_headless.libs"
OBLIB="$(ls "$OBDIR"/libopenblasp-*.so | head -1)"
CFLAGS="-O2 -std=gnu99 -DHEADERPATH -DUNDERBLAS -I$REF -I$ROOT/include -fPIC -w"
CFLAGS="$CFLAGS -I$HERE"
objcopy --redefine-sym HFpLinsysCreate=HFpLinsysCreate_ref "$OBJ/hdsdp_linsolver.o" "$OUT/hdsdp_linsolver_renamed.o"
objcopy --redefine-sym HConeSetData=HConeSetData_ref "$OBJ/hdsdp_conic.o" "$OUT/hdsdp_conic_renamed.o"
objcopy --redefine-sym fds_syev=fds_syev_dimacs "$OBJ/hdsdp.o" "$OUT/hdsdp_renamed.o"
HOOKS=""
for f in hdsdp_schur_cuda hdsdp_linsys_cuda hdsdp_conic_cuda hdsdpcu_shim; do
    gcc $CFLAGS -c "$HERE/$f.c" -o "$OUT/$f.o"
    HOOKS="$HOOKS $OUT/$f.o"
done
HOOKS="$HOOKS $OUT/hdsdp_linsolver_renamed.o $OUT/hdsdp_conic_renamed.o $OUT/hdsdp_renamed.o"
REFOBJS="$(ls "$OBJ"/*.o | grep -v -e /hdsdp_schur.o -e /hdsdp_linsolver.o -e /hdsdp_conic.o -e /hdsdp.o -e /ref_driver.o)"
LINK="-L$ROOT/hdsdp_b200 -lhdsdp_cuda -Wl,--disable-new-dtags,-rpath,\$ORIGIN/../../hdsdp_b200,-rpath,$OBDIR $OBLIB -lm"
gcc -shared -o "$OUT/libhdsdp_integrated.so" $REFOBJS "$OBJ/ref_driver.o" $HOOKS $LINK
gcc $CFLAGS "$REF/tests/sdpasolve.c" $REFOBJS $HOOKS $LINK -o "$OUT/sdpasolve_cuda"
echo "build_integrated: built $OUT/libhdsdp_integrated.so and $OUT/sdpasolve_cuda"
