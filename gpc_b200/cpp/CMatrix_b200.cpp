// CMatrix_b200.cpp -- INTEGRATION.md level 1 as a compiled object: the six CMatrix methods that wrap the hot LAPACK / BLAS
// calls of lapack.h, re-bound to the CMatrix-level C ABI of libgpc_b200.so (gpc_dpotrf / dpotri / dtrsm / dsyrk / dgemm /
// dsymv: host pointers, LAPACK argument meaning, staged through device memory).
//
// The reference's CMatrix.cpp defines the same member functions.  No source is changed: the build recipe WEAKENS those six
// symbols in the reference's own CMatrix.o
//     objcopy --weaken-symbol=_ZN7CMatrix5potrfEPKc ... CMatrix.o CMatrix_weak.o          (gpc_b200/cpp/Makefile: level1)
// and links this object next to it; the strong definitions below win, every other CMatrix method is the reference's.
// Argument checks and exceptions are the reference's (CMatrix.cpp:127-140, 205-247, 272-322, 371-379, 414-420).
// CMatrix::syr (dsyr_) is defined inline in CMatrix.h:526-533 and cannot be re-bound without touching the header; its
// drop-in gpc_dsyr is exported for a maintainer who moves it out of line.
#include <cstdlib>
#include "CMatrix.h"
#include "gpc_b200.h"

namespace
{
int gpcDevice()
{
  static int d = getenv("GPC_DEVICE") ? atoi(getenv("GPC_DEVICE")) : 0;
  return d;
}
void gpcFail() { throw ndlexceptions::Error(std::string("gpc_b200: ") + gpc_last_error()); }
} // namespace

void CMatrix::potrf(const char* type) // CMatrix.cpp:371-379, was dpotrf_
{
  MATRIXPROPERTIES(isSymmetric());
  int info = 0;
  if(gpc_dpotrf(gpcDevice(), type[0], nrows, vals, ncols, &info) < 0)
    gpcFail();
  setSymmetric(false);
  setTriangular(true);
  if(info != 0)
    throw ndlexceptions::MatrixNonPosDef(); // what jitChol catches (CMatrix.cpp:785-792)
}
void CMatrix::potri(const char* type) // CMatrix.cpp:414-420, was dpotri_
{
  MATRIXPROPERTIES(isSquare());
  int info = 0;
  if(gpc_dpotri(gpcDevice(), type[0], nrows, vals, ncols, &info) < 0)
    gpcFail();
  if(info != 0)
    throw ndlexceptions::MatrixNonPosDef();
}
void CMatrix::trsm(const CMatrix& A, double alpha, const char* side, const char* type, const char* trans, const char* diag)
{ // CMatrix.cpp:272-295, was dtrsm_: same argument checks, then the device call
  CHARARGUMENTS(side[0] == 'L' || side[0] == 'l' || side[0] == 'R' || side[0] == 'r');
  CHARARGUMENTS(type[0] == 'L' || type[0] == 'l' || type[0] == 'U' || type[0] == 'u');
  CHARARGUMENTS(trans[0] == 'N' || trans[0] == 'n' || trans[0] == 'T' || trans[0] == 't');
  CHARARGUMENTS(diag[0] == 'N' || diag[0] == 'n' || diag[0] == 'U' || diag[0] == 'u');
  MATRIXPROPERTIES(A.isTriangular());
  const bool left = (side[0] == 'L' || side[0] == 'l');
  DIMENSIONMATCH(A.nrows == (left ? nrows : ncols));
  if(gpc_dtrsm(gpcDevice(), side[0], type[0], trans[0], diag[0], nrows, ncols, alpha, A.vals, A.nrows, vals, nrows) < 0)
    gpcFail();
}
void CMatrix::syrk(const CMatrix& A, double alpha, double beta, const char* type, const char* trans)
{ // CMatrix.cpp:297-322, was dsyrk_: C := alpha op(A) op(A)' + beta C
  MATRIXPROPERTIES(isSymmetric() || beta == 0.0);
  CHARARGUMENTS(trans[0] == 'n' || trans[0] == 'N' || trans[0] == 't' || trans[0] == 'T');
  const bool notrans = (trans[0] == 'n' || trans[0] == 'N');
  const unsigned int n = notrans ? ncols : nrows, k = notrans ? A.ncols : A.nrows;
  DIMENSIONMATCH(n == (notrans ? A.nrows : A.ncols));
  if(gpc_dsyrk(gpcDevice(), type[0], trans[0], n, k, alpha, A.vals, A.nrows, beta, vals, nrows) < 0)
    gpcFail();
  copySymmetric(type);
}
void CMatrix::gemm(const CMatrix& A, const CMatrix& B, double alpha, double beta, const char* transa, const char* transb)
{ // CMatrix.cpp:205-247, was dgemm_
  setSymmetric(false);
  unsigned int m = 0, n = 0, k = 0;
  switch(transa[0])
  {
  case 'n':
  case 'N':
    m = A.nrows;
    k = A.ncols;
    break;
  case 't':
  case 'T':
    m = A.ncols;
    k = A.nrows;
    break;
  default:
    throw ndlexceptions::Error("No such value for transa.");
  }
  switch(transb[0])
  {
  case 'n':
  case 'N':
    n = B.ncols;
    DIMENSIONMATCH(k == B.nrows);
    break;
  case 't':
  case 'T':
    n = B.nrows;
    DIMENSIONMATCH(k == B.ncols);
    break;
  default:
    throw ndlexceptions::Error("No such value for transb.");
  }
  DIMENSIONMATCH(n == ncols);
  DIMENSIONMATCH(m == nrows);
  if(gpc_dgemm(gpcDevice(), transa[0], transb[0], m, n, k, alpha, A.vals, A.nrows, B.vals, B.nrows, beta, vals, nrows) < 0)
    gpcFail();
}
void CMatrix::symv(const CMatrix& A, const CMatrix& x, double alpha, double beta, const char* upperLower)
{ // CMatrix.cpp:127-140, was dsymv_
  MATRIXPROPERTIES(A.isSymmetric());
  DIMENSIONMATCH(ncols == 1);
  DIMENSIONMATCH(x.ncols == 1);
  CHARARGUMENTS(upperLower[0] == 'u' || upperLower[0] == 'U' || upperLower[0] == 'l' || upperLower[0] == 'L');
  DIMENSIONMATCH(nrows == A.nrows);
  DIMENSIONMATCH(nrows == x.nrows);
  if(gpc_dsymv(gpcDevice(), upperLower[0], A.ncols, alpha, A.vals, A.nrows, x.vals, beta, vals) < 0)
    gpcFail();
}
