#!/bin/bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Compiles the reference's own sources WHERE THEY LIE under /root/reference
# (nothing is copied into this repo) plus oracle/ref_driver.cpp into
# oracle/_ref/ref_driver.  Recipe follows SURVEY.md §8(c).  Needs only g++/gcc
# and the LP64 OpenBLAS that ships inside the opencv wheel of this image.
# On the GPU box /root/reference is absent: the prebuilt oracle/_ref/ travels
# with the snapshot and this script exits 0 without doing anything.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${LK_REFERENCE:-/root/reference}
OUT="$HERE/_ref"
if [ ! -d "$REF/src/lib" ]; then
  echo "[build_ref] $REF not present; keeping prebuilt $OUT" ; exit 0
fi
SP=$(python -c "import site;print(site.getsitepackages()[0])")
OB="$SP/opencv_python_headless.libs"
OBLIB=$(ls "$OB" | grep '^libopenblas' | head -1)
if [ -z "$OBLIB" ]; then echo "[build_ref] no LP64 OpenBLAS found under $OB"; exit 1; fi
mkdir -p "$OUT/obj"
if [ "$OUT/ref_driver" -nt "$HERE/ref_driver.cpp" ] && [ "$OUT/ref_driver" -nt "$HERE/build_ref.sh" ] && \
   [ "$OUT/ref_nested_driver" -nt "$HERE/ref_nested_driver.cpp" ] && [ "$OUT/ref_nested_driver" -nt "$HERE/build_ref.sh" ]; then
  echo "[build_ref] up to date"; exit 0
fi
cat > "$OUT/obj/version_stub.cpp" <<'EOS'
#include "libKriging/version.hpp"
std::string libKriging::version() { return "oracle-build"; }
std::string libKriging::buildTag() { return "oracle-build"; }
EOS
CXXFLAGS="-O2 -std=c++17 -fopenmp -fPIC -DARMA_DONT_USE_WRAPPER -DARMA_DONT_USE_OPENMP -DARMA_32BIT_WORD -DARMA_USE_BLAS -DARMA_USE_LAPACK -DNDEBUG"
INC="-I$REF/dependencies/armadillo-code/include -I$REF/src/lib/include -I$REF/src/lib -I$REF/dependencies/lbfgsb_cpp/include"
pids=()
for f in Kriging KrigingImpl LinearAlgebra Covariance Optim Trend Random Bench lkalloc utils/jsonutils utils/base64; do
  o="$OUT/obj/$(basename $f).o"
  g++ $CXXFLAGS $INC -c "$REF/src/lib/$f.cpp" -o "$o" &
  pids+=($!)
done
g++ $CXXFLAGS $INC -c "$OUT/obj/version_stub.cpp" -o "$OUT/obj/version_stub.o" &
pids+=($!)
g++ $CXXFLAGS $INC -c "$HERE/ref_driver.cpp" -o "$OUT/obj/ref_driver.o" &
pids+=($!)
# NestedKriging and what it links against, for ref_nested_driver
mkdir -p "$OUT/obj_nested"
for f in NestedKriging WarpKriging; do
  g++ $CXXFLAGS $INC -c "$REF/src/lib/$f.cpp" -o "$OUT/obj_nested/$f.o" &
  pids+=($!)
done
g++ $CXXFLAGS $INC -c "$HERE/ref_nested_driver.cpp" -o "$OUT/obj_nested/ref_nested_driver.o" &
pids+=($!)
for f in blas lbfgsb linpack s_cmp s_copy timer; do
  gcc -O2 -fPIC -w -Ddcopy_=Wcopy_ -Ddscal_=Wscal_ -Ddaxpy_=Waxpy_ -Ddnrm2_=Wnrm2_ -Dddot_=Wdot_ \
      -I"$REF/dependencies/lbfgsb_cpp/Lbfgsb.3.0" -I"$REF/dependencies/lbfgsb_cpp/Lbfgsb.3.0/include" \
      -c "$REF/dependencies/lbfgsb_cpp/Lbfgsb.3.0/$f.c" -o "$OUT/obj/lb_$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
objs=$(ls "$OUT"/obj/*.o | grep -v '/ref_driver.o$')
g++ -fopenmp -o "$OUT/ref_nested_driver" $objs "$OUT"/obj_nested/*.o -L"$OB" -l:"$OBLIB" -Wl,-rpath,"$OB" -Wl,-rpath,"$SP/scipy.libs" -lpthread
g++ -fopenmp -o "$OUT/ref_driver" "$OUT"/obj/*.o -L"$OB" -l:"$OBLIB" -Wl,-rpath,"$OB" -Wl,-rpath,"$SP/scipy.libs" -lpthread
echo "$OB:$SP/scipy.libs" > "$OUT/ld_library_path.txt"
echo "[build_ref] built $OUT/ref_driver"
