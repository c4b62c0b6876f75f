#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle): builds the reference's own C++ SSIDS CPU engine,
# unmodified, from the sources where they lie under /root/reference, into oracle/_ref/
# (git-ignored; travels to the GPU box with the snapshot).  Nothing is copied into the repo:
# the include tree under oracle/_ref/inc is made of symlinks plus a sed-generated config.h.
#
# Output: oracle/_ref/libssids_ref.so  (reference engine + our thin driver oracle/ref_driver.cpp)
#
# Reference sources compiled: src/ssids/cpu/{NumericSubtree,SymbolicSubtree,
# ThreadStats}.cxx, src/ssids/cpu/kernels/{cholesky,ldlt_app,ldlt_nopiv,ldlt_tpp,wrappers}.cxx,
# src/spral/{omp,compat}.cxx, src/ssids/profile.cxx.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GALAHAD_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/ssids/cpu" ]; then
  echo "build_ref: $REF not present; keeping any prebuilt $OUT/libssids_ref.so" >&2
  exit 0
fi
BLASDIR="$(python3 - <<'EOF'
import glob, sys, os, sysconfig
sp = sysconfig.get_paths()["purelib"]
c = glob.glob(os.path.join(sp, "opencv_python_headless.libs", "libopenblasp-*.so"))
print(os.path.dirname(c[0]) if c else "")
EOF
)"
if [ -z "$BLASDIR" ]; then echo "build_ref: no OpenBLAS with standard symbols found" >&2; exit 1; fi
BLASLIB="$(ls "$BLASDIR"/libopenblasp-*.so | head -1)"

mkdir -p "$OUT/inc/ssids" "$OUT/inc/hw_topology" "$OUT/obj"
ln -sfn "$REF/src/ssids/cpu" "$OUT/inc/ssids/cpu"
ln -sf "$REF/src/ssids/contrib.h" "$OUT/inc/ssids/contrib.h"
ln -sf "$REF/src/ssids/profile.hxx" "$OUT/inc/ssids/profile.hxx"
ln -sf "$REF/src/spral/omp.hxx" "$OUT/inc/omp.hxx"
ln -sf "$REF/src/spral/compat.hxx" "$OUT/inc/compat.hxx"
ln -sf "$REF/src/spral/guess_topology.hxx" "$OUT/inc/hw_topology/guess_topology.hxx"
ln -sf "$REF/src/spral/hwloc_wrapper.hxx" "$OUT/inc/hw_topology/hwloc_wrapper.hxx"
# config.h: the reference's, with hwloc switched off (not installed here)
sed -e 's|^#define HAVE_HWLOC 1|/* #undef HAVE_HWLOC */|' "$REF/src/spral/config.h" > "$OUT/inc/config.h"

CXX="${REF_CXX:-/usr/bin/g++}"
FLAGS="-std=c++11 -O2 -fopenmp -fPIC -I$OUT/inc"
SRCS=(
  "$REF/src/ssids/cpu/NumericSubtree.cxx"
  "$REF/src/ssids/cpu/SymbolicSubtree.cxx"
  "$REF/src/ssids/cpu/ThreadStats.cxx"
  "$REF/src/ssids/cpu/kernels/cholesky.cxx"
  "$REF/src/ssids/cpu/kernels/ldlt_app.cxx"
  "$REF/src/ssids/cpu/kernels/ldlt_nopiv.cxx"
  "$REF/src/ssids/cpu/kernels/ldlt_tpp.cxx"
  "$REF/src/ssids/cpu/kernels/wrappers.cxx"
  "$REF/src/spral/omp.cxx"
  "$REF/src/spral/compat.cxx"
  "$REF/src/ssids/profile.cxx"
)
OBJS=()
pids=()
for s in "${SRCS[@]}"; do
  o="$OUT/obj/$(basename "${s%.cxx}").o"
  OBJS+=("$o")
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ]; then
    $CXX $FLAGS -c "$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$CXX $FLAGS -c "$HERE/ref_driver.cpp" -o "$OUT/obj/ref_driver.o"
$CXX -shared -fopenmp -o "$OUT/libssids_ref.so" "${OBJS[@]}" "$OUT/obj/ref_driver.o" \
  "$BLASLIB" -Wl,--disable-new-dtags -Wl,-rpath,"$BLASDIR"
echo "build_ref: built $OUT/libssids_ref.so (BLAS: $BLASLIB)"
