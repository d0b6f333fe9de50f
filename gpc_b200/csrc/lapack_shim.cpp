// lapack_shim.cpp -- seam (1) of SURVEY.md 8(b): the Fortran BLAS/LAPACK symbols GPc declares in lapack.h:17-232, for
// the five calls that carry its hot path, forwarded to libgpc_b200.so.  Linking (or LD_PRELOADing) this in front of
// the BLAS lets the UNMODIFIED reference objects -- CMatrix::potrf/potri/trsm/syrk/gemm, lapack.h:59-73, 186-222 --
// run their Cholesky, inverse, triangular solves and products on the B200 without a source change (the level-1
// binding of INTEGRATION.md, obtained at link time).  Host pointers in, host pointers out; every call is staged
// through device memory, so this is the compatibility path, not the fast one (that is gpc_eval).
//
// Fortran ABI: every scalar by reference, flags as char* of which only [0] is read, column-major, 32-bit ints.
#include <stdio.h>
#include <stdlib.h>
#include "../../include/gpc_b200.h"

#if defined(__GNUC__)
#define SHIM_EXPORT __attribute__((visibility("default")))
#else
#define SHIM_EXPORT
#endif

static int shim_device() {
  static int dev = -1;
  if (dev < 0) {
    const char* e = getenv("GPC_DEVICE");
    dev = e ? atoi(e) : 0;
  }
  return dev;
}
static void shim_fail(const char* what) {
  // the Fortran interfaces have no error channel besides info: a CUDA failure is fatal, as a missing BLAS would be
  fprintf(stderr, "gpc_b200 lapack shim: %s failed: %s\n", what, gpc_last_error());
  abort();
}

extern "C" {

SHIM_EXPORT void dpotrf_(const char* uplo, const int* n, double* a, const int* lda, int* info) {
  *info = 0;
  if (*n <= 0) return;
  int rc = gpc_dpotrf(shim_device(), uplo[0], *n, a, *lda, info);
  if (rc < 0) shim_fail("dpotrf_");
}

SHIM_EXPORT void dpotri_(const char* uplo, const int* n, double* a, const int* lda, int* info) {
  *info = 0;
  if (*n <= 0) return;
  int rc = gpc_dpotri(shim_device(), uplo[0], *n, a, *lda, info);
  if (rc < 0) shim_fail("dpotri_");
}

SHIM_EXPORT void dtrsm_(const char* side, const char* uplo, const char* trans, const char* diag, const int* m,
                        const int* n, const double* alpha, const double* a, const int* lda, double* b, const int* ldb) {
  if (*m <= 0 || *n <= 0) return;
  int rc = gpc_dtrsm(shim_device(), side[0], uplo[0], trans[0], diag[0], *m, *n, *alpha, a, *lda, b, *ldb);
  if (rc < 0) shim_fail("dtrsm_");
}

SHIM_EXPORT void dsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha,
                        const double* a, const int* lda, const double* beta, double* c, const int* ldc) {
  if (*n <= 0) return;
  if (*k <= 0) {  // C := beta C on the referenced triangle
    const bool up = uplo[0] == 'U' || uplo[0] == 'u';
    for (int j = 0; j < *n; j++)
      for (int i = up ? 0 : j; i < (up ? j + 1 : *n); i++) c[i + (long)j * *ldc] = (*beta == 0.0) ? 0.0 : *beta * c[i + (long)j * *ldc];
    return;
  }
  int rc = gpc_dsyrk(shim_device(), uplo[0], trans[0], *n, *k, *alpha, a, *lda, *beta, c, *ldc);
  if (rc < 0) shim_fail("dsyrk_");
}

SHIM_EXPORT void dgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k,
                        const double* alpha, const double* a, const int* lda, const double* b, const int* ldb,
                        const double* beta, double* c, const int* ldc) {
  if (*m <= 0 || *n <= 0) return;
  if (*k <= 0) {
    for (int j = 0; j < *n; j++)
      for (int i = 0; i < *m; i++) c[i + (long)j * *ldc] = (*beta == 0.0) ? 0.0 : *beta * c[i + (long)j * *ldc];
    return;
  }
  int rc = gpc_dgemm(shim_device(), transa[0], transb[0], *m, *n, *k, *alpha, a, *lda, b, *ldb, *beta, c, *ldc);
  if (rc < 0) shim_fail("dgemm_");
}

}  // extern "C"
