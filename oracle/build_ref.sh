#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Compiles the UNMODIFIED reference (SheffieldML/GPc) from the sources WHERE THEY LIE under
# $GPC_REFERENCE (default /root/reference) into oracle/_ref/ (git-ignored, travels with gpurun):
#   oracle/_ref/libgpcref.so   reference objects + oracle/ref_driver.cpp (ctypes facade)
#   oracle/_ref/gp, gplvm, ivm the reference CLI front-ends (for end-to-end trajectories)
# No reference source is copied into this repository.  The reference's own build system is not used
# (it wants gfortran + system BLAS): three shims instead:
#   1. -std=gnu++98 (dynamic exception specs, ostream->void*)
#   2. ndlfortran.c (the shipped f2c translation) with oracle/shim/f2c.h; lbfgs_ stubbed
#   3. BLAS/LAPACK = the scipy wheel's LP64 OpenBLAS (symbols prefixed scipy_) via a generated blasmap.h
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GPC_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) - keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
PY="${PYTHON:-python3}"
SP=$($PY -c "import scipy,os;print(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)),'scipy.libs'))")
OB=$(ls "$SP"/libscipy_openblas-*.so | head -1)
# the 21 Fortran symbols declared in the reference's lapack.h:17-232
for s in dsyev dsysv dgetrf dgetri dpotrf dpotri dswap dcopy dscal daxpy ddot dnrm2 dgemv dsymv dger dsyr dgemm dsyrk dtrmm dtrsm dsymm; do
  echo "#define ${s}_ scipy_${s}_"
done > "$OUT/obj/blasmap.h"
CXXF="-std=gnu++98 -O3 -fPIC -w -D_LINUX -include $OUT/obj/blasmap.h -I$REF"
SRCS="CClctrl CGp CGplvm CIvm CMatrix CNoise ndlutil ndlstrutil CTransform COptimisable CKern CDist ndlassert gp gplvm ivm"
pids=()
for f in $SRCS; do
  if [ ! -f "$OUT/obj/$f.o" ] || [ "$REF/$f.cpp" -nt "$OUT/obj/$f.o" ]; then
    g++ $CXXF -c "$REF/$f.cpp" -o "$OUT/obj/$f.o" &
    pids+=($!)
  fi
done
gcc -O3 -fPIC -w -I"$HERE/shim" -I"$REF" -c "$REF/ndlfortran.c" -o "$OUT/obj/ndlfortran.o"
gcc -O2 -fPIC -c "$HERE/shim/lbfgs_stub.c" -o "$OUT/obj/lbfgs_stub.o"
g++ $CXXF -c "$HERE/ref_driver.cpp" -o "$OUT/obj/ref_driver.o"
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
cd "$OUT/obj"
COMMON="CClctrl.o CMatrix.o ndlfortran.o lbfgs_stub.o CNoise.o ndlutil.o ndlstrutil.o CTransform.o COptimisable.o CKern.o CDist.o ndlassert.o"
g++ -shared -o "$OUT/libgpcref.so" ref_driver.o CGp.o CGplvm.o $COMMON "$OB" -Wl,-rpath,"$SP" -lm
g++ -o "$OUT/gp" gp.o CGp.o $COMMON "$OB" -Wl,-rpath,"$SP" -lm
g++ -o "$OUT/gplvm" gplvm.o CGplvm.o $COMMON "$OB" -Wl,-rpath,"$SP" -lm
g++ -o "$OUT/ivm" ivm.o CIvm.o $COMMON "$OB" -Wl,-rpath,"$SP" -lm
# ---- the UNMODIFIED reference with its five hot BLAS/LAPACK calls resolved by the B200 library (seam (1)):
# same sources, same flags, but dpotrf_/dpotri_/dtrsm_/dsyrk_/dgemm_ keep their plain Fortran names and are taken from
# gpc_b200/libgpc_lapack_shim.so; everything else still comes from OpenBLAS.  Built only when the shim exists.
SHIMDIR="$(cd "$HERE/../gpc_b200" && pwd)"
if [ -f "$SHIMDIR/libgpc_lapack_shim.so" ]; then
  mkdir -p "$OUT/obj_b200"
  grep -v -E "define (dpotrf|dpotri|dtrsm|dsyrk|dgemm)_ " "$OUT/obj/blasmap.h" > "$OUT/obj_b200/blasmap.h"
  CXXB="-std=gnu++98 -O3 -fPIC -w -D_LINUX -include $OUT/obj_b200/blasmap.h -I$REF"
  pids=()
  for f in CClctrl CGp CGplvm CIvm CMatrix CNoise ndlutil ndlstrutil CTransform COptimisable CKern CDist ndlassert gp gplvm ivm; do
    if [ ! -f "$OUT/obj_b200/$f.o" ] || [ "$REF/$f.cpp" -nt "$OUT/obj_b200/$f.o" ]; then
      g++ $CXXB -c "$REF/$f.cpp" -o "$OUT/obj_b200/$f.o" &
      pids+=($!)
    fi
  done
  for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
  cd "$OUT/obj_b200"
  g++ -o "$OUT/gp_b200" gp.o CGp.o CClctrl.o CMatrix.o ../obj/ndlfortran.o ../obj/lbfgs_stub.o CNoise.o ndlutil.o ndlstrutil.o \
      CTransform.o COptimisable.o CKern.o CDist.o ndlassert.o -L"$SHIMDIR" -lgpc_lapack_shim -lgpc_b200 "$OB" \
      -Wl,-rpath,"$SP" -Wl,-rpath,'$ORIGIN/../../gpc_b200' -lm
  g++ -o "$OUT/gplvm_b200" gplvm.o CGplvm.o CClctrl.o CMatrix.o ../obj/ndlfortran.o ../obj/lbfgs_stub.o CNoise.o ndlutil.o ndlstrutil.o \
      CTransform.o COptimisable.o CKern.o CDist.o ndlassert.o -L"$SHIMDIR" -lgpc_lapack_shim -lgpc_b200 "$OB" \
      -Wl,-rpath,"$SP" -Wl,-rpath,'$ORIGIN/../../gpc_b200' -lm
  g++ -o "$OUT/ivm_b200" ivm.o CIvm.o CClctrl.o CMatrix.o ../obj/ndlfortran.o ../obj/lbfgs_stub.o CNoise.o ndlutil.o ndlstrutil.o \
      CTransform.o COptimisable.o CKern.o CDist.o ndlassert.o -L"$SHIMDIR" -lgpc_lapack_shim -lgpc_b200 "$OB" \
      -Wl,-rpath,"$SP" -Wl,-rpath,'$ORIGIN/../../gpc_b200' -lm
  echo "build_ref: built $OUT/gp_b200, gplvm_b200, ivm_b200 (reference objects + gpc_b200 Fortran shim)"
fi
# ---- level 2: the reference's UNMODIFIED front-ends compiled on the device-backed model classes of
# gpc_b200/cpp (CGpB200 : CGp, CGplvmB200 : CGplvm) through a prefix header -- `-include gp_dropin.h` redirects the names
# CGp / readGpFromFile inside gp.cpp only.  Every other object is the plain reference build of obj/ (host OpenBLAS for
# whatever is not the hot path); the hot path goes through the C ABI of libgpc_b200.so.  Also the in-process comparison
# driver of tests/cpp (reference class vs drop-in class on the same data).
CPPDIR="$SHIMDIR/cpp"
if [ -f "$SHIMDIR/libgpc_b200.so" ] && [ -d "$CPPDIR" ]; then
  mkdir -p "$OUT/obj_l2"
  CXXL="-std=gnu++98 -O3 -fPIC -w -D_LINUX -include $OUT/obj/blasmap.h -I$REF -I$HERE/../include -I$CPPDIR"
  pids=()
  for f in GpcKernBridge CGpB200 CGplvmB200 CCmpndKernB200; do
    g++ $CXXL -c "$CPPDIR/$f.cpp" -o "$OUT/obj_l2/$f.o" &
    pids+=($!)
  done
  g++ $CXXL -include "$CPPDIR/gp_dropin.h" -c "$REF/gp.cpp" -o "$OUT/obj_l2/gp.o" &
  pids+=($!)
  g++ $CXXL -include "$CPPDIR/gplvm_dropin.h" -c "$REF/gplvm.cpp" -o "$OUT/obj_l2/gplvm.o" &
  pids+=($!)
  g++ $CXXL -include "$CPPDIR/ivm_dropin.h" -c "$REF/ivm.cpp" -o "$OUT/obj_l2/ivm.o" &
  pids+=($!)
  g++ $CXXL -c "$HERE/../tests/cpp/cgp_b200_check.cpp" -o "$OUT/obj_l2/cgp_b200_check.o" &
  pids+=($!)
  for p in "${pids[@]}"; do wait "$p"; done
  cd "$OUT/obj"
  L2="../obj_l2/GpcKernBridge.o ../obj_l2/CGpB200.o ../obj_l2/CGplvmB200.o ../obj_l2/CCmpndKernB200.o"
  LNK="-L$SHIMDIR -lgpc_b200 $OB -Wl,-rpath,$SP -Wl,-rpath,\$ORIGIN/../../gpc_b200 -lm"
  g++ -o "$OUT/gp_l2" ../obj_l2/gp.o $L2 CGp.o CGplvm.o $COMMON $LNK
  g++ -o "$OUT/gplvm_l2" ../obj_l2/gplvm.o $L2 CGp.o CGplvm.o $COMMON $LNK
  g++ -o "$OUT/cgp_b200_check" ../obj_l2/cgp_b200_check.o $L2 CGp.o CGplvm.o $COMMON $LNK
  # the reference's ivm.cpp on CCmpndKernB200 (its kernel matrices K(X, X2) built on the device; `-include ivm_dropin.h`)
  g++ -o "$OUT/ivm_l2" ../obj_l2/ivm.o $L2 CIvm.o CGp.o CGplvm.o $COMMON $LNK
  # ---- level 1: the six CMatrix methods that wrap the hot lapack.h calls re-bound as a compiled object.
  # The reference's own CMatrix.o keeps every other method; its six definitions are weakened (objcopy, no source change)
  # so that the strong ones of gpc_b200/cpp/CMatrix_b200.cpp win at link time.
  mkdir -p "$OUT/obj_l1"
  objcopy --weaken-symbol=_ZN7CMatrix5potrfEPKc --weaken-symbol=_ZN7CMatrix5potriEPKc \
          --weaken-symbol=_ZN7CMatrix4trsmERKS_dPKcS3_S3_S3_ --weaken-symbol=_ZN7CMatrix4syrkERKS_ddPKcS3_ \
          --weaken-symbol=_ZN7CMatrix4gemmERKS_S1_ddPKcS3_ --weaken-symbol=_ZN7CMatrix4symvERKS_S1_ddPKc \
          CMatrix.o "$OUT/obj_l1/CMatrix_weak.o"
  g++ $CXXL -c "$CPPDIR/CMatrix_b200.cpp" -o "$OUT/obj_l1/CMatrix_b200.o"
  COMMON_L1="CClctrl.o ../obj_l1/CMatrix_b200.o ../obj_l1/CMatrix_weak.o ndlfortran.o lbfgs_stub.o CNoise.o ndlutil.o ndlstrutil.o CTransform.o COptimisable.o CKern.o CDist.o ndlassert.o"
  g++ -o "$OUT/gp_l1" gp.o CGp.o $COMMON_L1 $LNK
  g++ -o "$OUT/ivm_l1" ivm.o CIvm.o $COMMON_L1 $LNK
  # a host test double of libgpc_b200.so made of the reference's own classes (tests/cpp/mock_gpc_b200.cpp): lets the CPU
  # tests drive the DEVICE-path logic of the C++ host classes without a GPU (LD_LIBRARY_PATH=oracle/_ref/mock)
  mkdir -p "$OUT/mock"
  g++ $CXXL -c "$HERE/../tests/cpp/mock_gpc_b200.cpp" -o "$OUT/obj_l2/mock_gpc_b200.o"
  g++ -shared -o "$OUT/mock/libgpc_b200.so" "$OUT/obj_l2/mock_gpc_b200.o" $COMMON "$OB" -Wl,-rpath,"$SP" -lm
  echo "build_ref: built $OUT/gp_l2, gplvm_l2, ivm_l2, cgp_b200_check (reference front-ends on CGpB200 / CGplvmB200 / CCmpndKernB200), gp_l1, ivm_l1 (CMatrix level 1)"
fi
echo "build_ref: built $OUT/libgpcref.so, gp, gplvm (OpenBLAS: $OB)"
