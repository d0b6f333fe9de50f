// api_dev.cu -- device-level C ABI: the same kernels on CALLER-OWNED device memory and stream.  This is what the
// multi-GPU orchestration (gpc_b200/dist.py) drives: PyTorch owns device memory, streams and the NCCL collectives
// (plumbing); every flop still runs in the hand-written kernels of this library.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"

using namespace gpc;

#define GPC_CHECK(expr)            \
  do {                             \
    int _rc = (expr);              \
    if (_rc != GPC_OK) return _rc; \
  } while (0)

struct gpc_dev {
  int device;
  cudaStream_t stream;
  int64_t launches;
  double* partial;  // gradient partial sums
  double* gscr;     // reduced gradient (device)
  int max_ctas;
};

static Dense dense_of(gpc_dev* h, double* Dinv, int* info, double* logdet, int64_t nvalid) {
  Dense d;
  d.s = h->stream;
  d.launches = &h->launches;
  d.Dinv = Dinv;
  d.info = info;
  d.logdet = logdet;
  d.W = nullptr;
  d.nvalid = nvalid;
  return d;
}

extern "C" {

int gpc_dev_create(gpc_dev** out, int device, void* stream) {
  if (!out) return GPC_ERR_ARG;
  int ndev = 0;
  GPC_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) {
    set_error("gpc_dev_create: no such CUDA device");
    return GPC_ERR_CUDA;
  }
  GPC_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  GPC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  gpc_dev* h = new gpc_dev();
  memset(h, 0, sizeof(*h));
  h->device = device;
  h->stream = (cudaStream_t)stream;
  h->max_ctas = prop.multiProcessorCount * 2;
  GPC_CUDA_CHECK(cudaMalloc(&h->partial, (size_t)h->max_ctas * GPC_MAX_PARAMS * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&h->gscr, GPC_MAX_PARAMS * sizeof(double)));
  *out = h;
  return GPC_OK;
}
int gpc_dev_destroy(gpc_dev* h) {
  if (!h) return GPC_OK;
  cudaSetDevice(h->device);
  cudaFree(h->partial);
  cudaFree(h->gscr);
  delete h;
  return GPC_OK;
}
int gpc_dev_set_stream(gpc_dev* h, void* stream) {
  if (!h) return GPC_ERR_ARG;
  h->stream = (cudaStream_t)stream;
  return GPC_OK;
}
int64_t gpc_dev_launch_count(gpc_dev* h) { return h ? h->launches : 0; }

// in-place lower Cholesky of the n x n block at A (n multiple of 128).  Dinv: n x 128 doubles (inverses of the diagonal
// 128-blocks).  base = global index of the block's first row (for info), nvalid = number of real (unpadded) rows of
// the whole matrix.  info_dev / logdet_dev are device scalars that accumulate (zero them before the factorisation).
int gpc_dev_potrf(gpc_dev* h, double* A, int64_t lda, int64_t n, int64_t base, int64_t nvalid, double* Dinv,
                  int* info_dev, double* logdet_dev) {
  if (!h || !A || n < TILE || n % TILE || lda < n) {
    set_error("gpc_dev_potrf: bad arguments");
    return GPC_ERR_ARG;
  }
  GPC_CUDA_CHECK(cudaSetDevice(h->device));
  // the recursion indexes Dinv / info by global block position: shift so that local block 0 maps to Dinv[0]
  Dense d = dense_of(h, Dinv - base * TILE, info_dev, logdet_dev, nvalid);
  return potrf_rec(d, A, lda, n, base, nullptr);
}
// trans = 'T': X L' = B;  'N': X L = B.  B (m x n) in place, L n x n lower with its Dinv (n x 128).
int gpc_dev_trsm(gpc_dev* h, char trans, double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl, int64_t n,
                 const double* Dinv) {
  if (!h || !B || !L || !Dinv || m % TILE || n % TILE || m < 0 || n < TILE) {
    set_error("gpc_dev_trsm: bad arguments");
    return GPC_ERR_ARG;
  }
  GPC_CUDA_CHECK(cudaSetDevice(h->device));
  Dense d = dense_of(h, const_cast<double*>(Dinv), nullptr, nullptr, n);
  if (trans == 'T' || trans == 't') return trsm_rlt(d, B, ldb, m, L, ldl, n, 0);
  return trsm_rln(d, B, ldb, m, L, ldl, n, 0);
}
int gpc_dev_gemm(gpc_dev* h, int a_kc, int b_kc, int lower, int64_t m, int64_t n, int64_t k, double alpha,
                 const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
  if (!h) return GPC_ERR_ARG;
  GPC_CUDA_CHECK(cudaSetDevice(h->device));
  // lower: bit 0 = lower-triangle tiles only; bit 1 = A operand zero for kk < i (k loop of row tile r0 starts at r0)
  GemmCall g{A, B, C, lda, ldb, ldc, m, n, k, alpha, beta, a_kc != 0, b_kc != 0, (lower & 1) != 0};
  g.ktri = (lower & 2) != 0;
  return launch_gemm(g, h->stream, &h->launches);
}
// columns [col0, col0+ncols) of the training kernel matrix of X (n real rows, np padded rows), all np rows,
// written to K + col0*ldk (i.e. K addresses the full matrix)
int gpc_dev_kbuild_cols(gpc_dev* h, const gpc_kcomp* comps, int ncomp, const double* X, int64_t ldx, int64_t n,
                        int64_t np, int D, int64_t col0, int64_t ncols, double* K, int64_t ldk) {
  if (!h || !X || !K || col0 % 64 || ncols % 64 || np % 64) {
    set_error("gpc_dev_kbuild_cols: bad arguments");
    return GPC_ERR_ARG;
  }
  GPC_CUDA_CHECK(cudaSetDevice(h->device));
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, D, &ks));
  int64_t n2 = n - col0 < 0 ? 0 : (n - col0 < ncols ? n - col0 : ncols);
  return launch_kcross(ks, X, ldx, n, np, X + col0, ldx, n2, ncols, K + col0 * ldk, ldk, h->stream, &h->launches, col0);
}
// gradient partial sums over the lower-triangle tiles of columns [col0, col0+ncols): Cg addresses the full matrix
// (Cg[i + j*ldc] must be valid for j in the column range, i >= j); result (nparams doubles) -> host g_out after a sync
int gpc_dev_grad_cols(gpc_dev* h, const gpc_kcomp* comps, int ncomp, const double* X, int64_t ldx, int64_t n, int D,
                      int64_t col0, int64_t ncols, const double* Cg, int64_t ldc, const double* alpha, int64_t lda,
                      int dout, double* g_out) {
  if (!h || !g_out) return GPC_ERR_ARG;
  GPC_CUDA_CHECK(cudaSetDevice(h->device));
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, D, &ks));
  GPC_CHECK(launch_grad(ks, X, ldx, n, round_up(n, TILE), Cg, ldc, alpha, lda, dout, 0, h->partial, h->max_ctas,
                        h->gscr, nullptr, 0, h->stream, &h->launches, col0, ncols));
  GPC_CUDA_CHECK(cudaMemcpyAsync(g_out, h->gscr, ks.nparams * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  GPC_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return GPC_OK;
}

}  // extern "C"
