#!/bin/bash
# Builds the C++ host (lkgpu::Kriging + its command-line driver) against liblkgpu.so.
# Third-party host dependencies are the reference's own, used from where the reference vendors them
# (nothing is copied into this repo): Armadillo (header-only here: ARMA_DONT_USE_BLAS/LAPACK/WRAPPER) and
# lbfgsb_cpp (header + the f2c'd Lbfgsb.3.0 C sources).  On a box without that tree (the GPU box) the prebuilt
# libkriging_b200/host/_build/ travels with the snapshot and this script exits 0.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
DEPS=${LK_DEPENDENCIES:-/root/reference/dependencies}
OUT="$HERE/_build"
if [ ! -d "$DEPS/armadillo-code/include" ] || [ ! -d "$DEPS/lbfgsb_cpp/include" ]; then
  echo "[build_host] $DEPS not present; keeping prebuilt $OUT"; exit 0
fi
LIB="$HERE/../liblkgpu.so"
if [ ! -f "$LIB" ]; then echo "[build_host] $LIB missing: run python -m libkriging_b200.build first"; exit 1; fi
mkdir -p "$OUT/obj"
newest=$(ls -t "$HERE"/*.cpp "$HERE"/*.hpp "$HERE"/build_host.sh "$HERE/../../include/lkgpu.h" | head -1)
if [ -x "$OUT/lkgpu_host_driver" ] && [ "$OUT/lkgpu_host_driver" -nt "$newest" ]; then echo "[build_host] up to date"; exit 0; fi
CXXFLAGS="-O2 -std=c++17 -fPIC -DARMA_DONT_USE_WRAPPER -DARMA_DONT_USE_BLAS -DARMA_DONT_USE_LAPACK -DARMA_DONT_USE_OPENMP -DNDEBUG"
INC="-I$DEPS/armadillo-code/include -I$DEPS/lbfgsb_cpp/include -I$HERE/../../include"
pids=()
g++ $CXXFLAGS $INC -c "$HERE/lkgpu_kriging.cpp" -o "$OUT/obj/lkgpu_kriging.o" & pids+=($!)
g++ $CXXFLAGS $INC -c "$HERE/lkgpu_host_driver.cpp" -o "$OUT/obj/lkgpu_host_driver.o" & pids+=($!)
g++ $CXXFLAGS -c "$HERE/lkgpu_comm.cpp" -o "$OUT/obj/lkgpu_comm.o" & pids+=($!)
for f in blas lbfgsb linpack s_cmp s_copy timer; do
  gcc -O2 -fPIC -w -I"$DEPS/lbfgsb_cpp/Lbfgsb.3.0" -I"$DEPS/lbfgsb_cpp/Lbfgsb.3.0/include" \
      -c "$DEPS/lbfgsb_cpp/Lbfgsb.3.0/$f.c" -o "$OUT/obj/lb_$f.o" & pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
g++ -shared -o "$OUT/liblkgpu_host.so" "$OUT/obj/lkgpu_kriging.o" "$OUT/obj/lkgpu_comm.o" "$OUT"/obj/lb_*.o -L"$HERE/.." -llkgpu -Wl,-rpath,'$ORIGIN/../..'
g++ -o "$OUT/lkgpu_host_driver" "$OUT/obj/lkgpu_host_driver.o" -L"$OUT" -llkgpu_host -L"$HERE/.." -llkgpu \
    -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../..' -lpthread
g++ -O2 -std=c++17 -o "$OUT/lkgpu_comm_selftest" "$HERE/lkgpu_comm_selftest.cpp" "$HERE/lkgpu_comm.cpp" -lpthread
echo "[build_host] built $OUT/lkgpu_host_driver"
