// sparse.cu -- the reference's sparse approximations of CGp on the device: DTC, DTCVAR and FITC
// with M inducing inputs X_u (CGp.cpp:713-735 K_uu / K_uf build, :766-861 updateAD, :939-988 logLikelihood,
// :1146-1223 / :1244-1413 gradients, :490-521 / :584-599 posterior).
//
// All three are the Gaussian log-density of the targets under Sigma = Q + Lambda, Q = K_fu K_uu^-1 K_uf,
//   DTC, DTCVAR: Lambda = I / beta;      FITC: Lambda = diag(k_ii - q_ii) + I / beta,
// evaluated in Woodbury form: with A = K_uu + K_uf Lambda^-1 K_fu (M x M; beta times the reference's A, CGp.cpp:770-772)
//   log|Sigma| = sum log lambda_i - log|K_uu| + log|A|,      Sigma^-1 m = Lambda^-1 m - (K_uf Lambda^-1)' A^-1 (K_uf Lambda^-1 m).
// The gradient is taken once from the density instead of following the reference's term-by-term code: with
// G = dL/dSigma = -1/2 sum_j (Sigma^-1 - a_j a_j'), a = Sigma^-1 m, B = K_uu^-1 K_uf, BS = B Sigma^-1 = A^-1 K_uf Lambda^-1:
//   dL/dK_uf = 2 B H,   dL/dK_uu = -B H B',   dL/d diag(K) = h,   dL/dbeta = -tr(G) / beta^2
//   DTC: H = G, h = 0;   FITC: H = G - diag(G), h = diag(G);   DTCVAR: DTC + the trace penalty -1/2 d beta tr(K - Q)
// and nothing N x N is ever formed: only diag(G), B G (M x N) and B G B' = -1/2 (d (K_uu^-1 - A^-1) - Ba Ba') (M x M).
// What runs where: K_uu (kbuild_kernel), K_uf (kcross_kernel), the two M x M factorisations with their explicit inverses
// (potrf_inv_rec / inverse_from_W), every M x M x N product (launch_gemm: the tensor-core engine for large N), the
// kernel-parameter and inducing-input gradients (grad_kernel: symmetric mode over K_uu, cross mode over K_uf), and the
// element-wise glue below.  The N-vectors (lambda, diag(G), k_ii - q_ii) are reduced on the host (3 N doubles).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"

#define GPC_CHECK(expr)            \
  do {                             \
    int _rc = (expr);              \
    if (_rc != GPC_OK) return _rc; \
  } while (0)

namespace gpc {

// ---- element-wise / reduction glue --------------------------------------------------------------------------------
// dst(:, i) = src(:, i) * w[i]
__global__ void sp_colscale_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t rows, int64_t cols,
                                   const double* __restrict__ w, double add) {
  const int64_t i = blockIdx.x;  // columns on grid.x: no 65535 limit
  const double wi = w[i] + add;
  for (int64_t r = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.y * blockDim.x)
    dst[r + i * rows] = src[r + i * rows] * wi;
}
// out[i] = sum_r A[r, i] B[r, i]   (one warp per column)
__global__ void sp_coldot_kernel(double* __restrict__ out, const double* __restrict__ A, const double* __restrict__ B,
                                 int64_t rows, int64_t cols) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= cols) return;
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int64_t r = lane; r < rows; r += 32) acc = fma(A[r + i * rows], B[r + i * rows], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[i] = acc;
}
// lambda_i and 1/lambda_i (0 in the padding)
__global__ void sp_lambda_kernel(double* __restrict__ lam, double* __restrict__ linv, const double* __restrict__ kdiag,
                                 const double* __restrict__ q, double beta, int fitc, int64_t n, int64_t np) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  if (i < n) {
    double l = 1.0 / beta;
    if (fitc) l += kdiag[i] - q[i];
    lam[i] = l;
    linv[i] = 1.0 / l;
  } else {
    lam[i] = 1.0;
    linv[i] = 0.0;
  }
}
// y(rows x d) += A(rows x cols) x(cols x d): one thread per row, the columns split over blockIdx.y, atomics into y
__global__ void sp_matvec_rows_kernel(const double* __restrict__ A, int64_t rows, int64_t cols, const double* __restrict__ x,
                                      int64_t ldx, int d, double* __restrict__ y, int64_t ldy, int64_t chunk) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int64_t c0 = (int64_t)blockIdx.y * chunk, c1 = (c0 + chunk < cols) ? c0 + chunk : cols;
  for (int o = 0; o < d; o++) {
    double acc = 0.0;
    for (int64_t c = c0; c < c1; c++) acc = fma(A[r + c * rows], x[c + (int64_t)o * ldx], acc);
    atomicAdd(&y[r + (int64_t)o * ldy], acc);
  }
}
// y(cols x d) = A(rows x cols)' x(rows x d): one warp per column
__global__ void sp_matvec_cols_kernel(const double* __restrict__ A, int64_t rows, int64_t cols, const double* __restrict__ x,
                                      int64_t ldx, int d, double* __restrict__ y, int64_t ldy) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= cols) return;
  const int lane = threadIdx.x & 31;
  for (int o = 0; o < d; o++) {
    double acc = 0.0;
    for (int64_t r = lane; r < rows; r += 32) acc = fma(A[r + i * rows], x[r + (int64_t)o * ldx], acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) y[i + (int64_t)o * ldy] = acc;
  }
}
// a = Lambda^-1 m - t  (N x d)
__global__ void sp_a_kernel(double* __restrict__ a, const double* __restrict__ m, const double* __restrict__ t,
                            const double* __restrict__ linv, int64_t np, int d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  for (int o = 0; o < d; o++) a[i + (int64_t)o * np] = linv[i] * m[i + (int64_t)o * np] - t[i + (int64_t)o * np];
}
// diag(G)_i = -1/2 (d (1/lambda_i - cd_i) - sum_o a_io^2),  cd = colsum(K_uf Lambda^-1 o BS);  h = (fitc ? diag(G) : 0) + c
__global__ void sp_gd_kernel(double* __restrict__ gd, double* __restrict__ h, const double* __restrict__ linv,
                             const double* __restrict__ cd, const double* __restrict__ a, int64_t n, int64_t np, int d,
                             int fitc, double c) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  double g = 0.0, hh = 0.0;
  if (i < n) {
    double aa = 0.0;
    for (int o = 0; o < d; o++) aa = fma(a[i + (int64_t)o * np], a[i + (int64_t)o * np], aa);
    g = -0.5 * ((double)d * (linv[i] - cd[i]) - aa);
    hh = (fitc ? g : 0.0) + c;
  }
  gd[i] = g;
  h[i] = hh;
}
// dL/dK_uf = 2 (BG - fitc B diag(gd)) - 2 c B,  BG = -1/2 (d BS - Ba a')
__global__ void sp_gkuf_kernel(double* __restrict__ out, const double* __restrict__ BS, const double* __restrict__ B,
                               const double* __restrict__ Ba, const double* __restrict__ a, const double* __restrict__ gd,
                               int64_t mp, int64_t np, int d, int fitc, double c) {
  const int64_t i = blockIdx.x;
  double ai[16];
  for (int o = 0; o < d && o < 16; o++) ai[o] = a[i + (int64_t)o * np];
  const double gi = fitc ? gd[i] : 0.0;
  for (int64_t r = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; r < mp; r += (int64_t)gridDim.y * blockDim.x) {
    double ba = 0.0;
    for (int o = 0; o < d; o++) ba = fma(Ba[r + (int64_t)o * mp], o < 16 ? ai[o] : a[i + (int64_t)o * np], ba);
    const double bg = -0.5 * ((double)d * BS[r + i * mp] - ba);
    const double b = B[r + i * mp];
    out[r + i * mp] = 2.0 * (bg - b * gi) - 2.0 * c * b;
  }
}
// dL/dK_uu (before the B Z B' term) = 1/2 (d (K_uu^-1 - A^-1) - Ba Ba')
__global__ void sp_gkuu_kernel(double* __restrict__ out, const double* __restrict__ Kuuinv, const double* __restrict__ Ainv,
                               const double* __restrict__ Ba, int64_t mp, int d) {
  const int64_t c = blockIdx.x;
  for (int64_t r = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; r < mp; r += (int64_t)gridDim.y * blockDim.x) {
    double bb = 0.0;
    for (int o = 0; o < d; o++) bb = fma(Ba[r + (int64_t)o * mp], Ba[c + (int64_t)o * mp], bb);
    out[r + c * mp] = 0.5 * ((double)d * (Kuuinv[r + c * mp] - Ainv[r + c * mp]) - bb);
  }
}
__global__ void sp_sub_kernel(double* __restrict__ out, const double* __restrict__ A, const double* __restrict__ B, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = A[i] - B[i];
}
// g[p] += sum_i h_i d k(x_i, x_i) / d theta_p  (diagComputeElement: CKern.cpp:165-171 and the per-kernel diag rules)
__global__ void sp_kdiag_grad_kernel(const __grid_constant__ KSpec ks, const double* __restrict__ X, int64_t ldx, int64_t n,
                                     const double* __restrict__ h, double* __restrict__ g) {
  __shared__ double sacc[GPC_MAX_PARAMS];
  for (int p = threadIdx.x; p < ks.nparams; p += blockDim.x) sacc[p] = 0.0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double hi = h[i];
    if (hi == 0.0) continue;
    double nrm = 0.0;
    if (ks.need_dot)
      for (int k = 0; k < ks.D; k++) {
        const double x = X[i + (int64_t)k * ldx];
        nrm = fma(x, x, nrm);
      }
    for (int c = 0; c < ks.ncomp; c++) {
      const double* p = ks.p + ks.poff[c];
      const int o = ks.poff[c];
      switch (ks.type[c]) {
        case GPC_KERN_WHITE:
        case GPC_KERN_BIAS: atomicAdd(&sacc[o], hi); break;
        case GPC_KERN_RBF:
        case GPC_KERN_RBFARD:
        case GPC_KERN_MATERN32:
        case GPC_KERN_MATERN52: atomicAdd(&sacc[o + 1], hi); break;  // k_ii = variance
        case GPC_KERN_LIN: atomicAdd(&sacc[o], hi * nrm); break;
        case GPC_KERN_POLY: {
          const double deg = ks.degree[c];
          const double arg = p[0] * nrm + p[1];
          const double pm1 = (deg == 2.0) ? arg : pow(arg, deg - 1.0);
          atomicAdd(&sacc[o], hi * p[2] * deg * pm1 * nrm);
          atomicAdd(&sacc[o + 1], hi * p[2] * deg * pm1);
          atomicAdd(&sacc[o + 2], hi * pm1 * arg);
        } break;
      }
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < ks.nparams; p += blockDim.x)
    if (sacc[p] != 0.0) atomicAdd(&g[p], sacc[p]);
}
// var[i] = kss[i] - sum_m T[i, m] Ksu[i, m] + 1/beta   (thread per row)
__global__ void sp_var_kernel(double* __restrict__ var, const double* __restrict__ kss, const double* __restrict__ T,
                              const double* __restrict__ Ksu, int64_t rows, int64_t ld, int64_t cols, double ibeta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double acc = 0.0;
  for (int64_t c = 0; c < cols; c++) acc = fma(T[i + c * ld], Ksu[i + c * ld], acc);
  var[i] = kss[i] - acc + ibeta;
}

}  // namespace gpc

using namespace gpc;

enum { SP_DTC = 1, SP_FITC = 2, SP_DTCVAR = 4 };  // CGp::DTC, FITC, DTCVAR (CGp.h:13-19)

struct gpc_sparse {
  int device, approx;
  int64_t N, Np;
  int M, Mp, D, d;
  cudaStream_t s;
  int64_t launches;
  double *X, *Mt, *Xu;
  double *Kuu, *Lu, *Wu, *Kuuinv, *A, *LA, *WA, *Ainv, *gKuu;  // Mp x Mp
  double *Kuf, *V, *B, *KufL, *BS, *gKuf;                      // Mp x Np
  double *kdiag, *q, *lam, *linv, *gd, *h, *cd;                // Np
  double *a, *ta;                                              // Np x d
  double *t1, *t2, *Ba;                                        // Mp x d
  double *tmpL, *Tpool, *symm_part, *partial, *gXu, *scal;
  int* info;
  double* hbuf;  // pinned: 4 Np + scalars
  int max_ctas;
  double beta;
  bool haveData, haveEval;
  // posterior scratch
  double *Xs, *Ksu, *Tq, *kss, *pmu;
  int64_t ps_cap;
};

static const int SS_LOGDET = 0, SS_QUAD = 1, SS_G0 = 8;  // scal: logdet accumulator, quad, then 3 gradient vectors

static Dense sp_dense(gpc_sparse* h, double* W) {
  Dense d;
  d.s = h->s;
  d.launches = &h->launches;
  d.Dinv = nullptr;
  d.info = h->info;
  d.logdet = h->scal + SS_LOGDET;
  d.W = nullptr;
  d.nvalid = h->M;
  d.Winv = W;
  d.ldw = h->Mp;
  d.tmpL = h->tmpL;
  d.Tpool = h->Tpool;
  d.TLpool = nullptr;
  return d;
}

// L L' = S (lower of S read), W = L^-1, *logdet = log|S|; the jitChol schedule on failure (CMatrix.cpp:767-804: S mutated)
static int sp_factor(gpc_sparse* h, double* S, double* L, double* W, double* logdet, double* jitter_added) {
  const int64_t Mp = h->Mp;
  double jitter = 0.0;
  *jitter_added = 0.0;
  for (int tries = 0;; tries++) {
    GPC_CUDA_CHECK(cudaMemsetAsync(h->info, 0, sizeof(int), h->s));
    GPC_CUDA_CHECK(cudaMemsetAsync(h->scal + SS_LOGDET, 0, sizeof(double), h->s));
    GPC_CHECK(launch_copy_lower(S, Mp, L, Mp, Mp, h->s, &h->launches));
    Dense d = sp_dense(h, W);
    GPC_CHECK(potrf_inv_rec(d, L, Mp, Mp, 0, d.Tpool, false, 0));
    int info = 0;
    GPC_CUDA_CHECK(cudaMemcpyAsync(&info, h->info, sizeof(int), cudaMemcpyDeviceToHost, h->s));
    GPC_CUDA_CHECK(cudaMemcpyAsync(logdet, h->scal + SS_LOGDET, sizeof(double), cudaMemcpyDeviceToHost, h->s));
    GPC_CUDA_CHECK(cudaStreamSynchronize(h->s));
    if (info == 0) return GPC_OK;
    if (tries == 0) {
      std::vector<double> dg((size_t)h->M);
      GPC_CUDA_CHECK(cudaMemcpy2D(dg.data(), sizeof(double), S, (Mp + 1) * sizeof(double), sizeof(double), h->M,
                                  cudaMemcpyDeviceToHost));
      double tr = 0.0;
      for (double v : dg) tr += v;
      jitter = 1e-6 * tr / (double)h->M;
    }
    GPC_CHECK(launch_add_diag(S, Mp, h->M, jitter, h->s, &h->launches));
    *jitter_added += jitter;
    jitter *= 10.0;
    if (jitter > 10.0 || tries + 1 >= 20) {
      set_error("sparse GP: matrix is non positive definite after jitter retries");
      return info;
    }
  }
}

// Out = S^-1 K for the factored S (W = L^-1): Out = W' (W K), V is scratch; K, V, Out are Mp x Np
static int sp_solve(gpc_sparse* h, const double* W, const double* K, double* V, double* Out) {
  const int64_t Mp = h->Mp, Np = h->Np;
  GemmCall g1{W, K, V, Mp, Mp, Mp, Mp, Np, Mp, 1.0, 0.0, false, true, false};
  g1.a_tri = -1;  // W lower: W(i, kk) = 0 for kk > i
  GPC_CHECK(launch_gemm(g1, h->s, &h->launches));
  GemmCall g2{W, V, Out, Mp, Mp, Mp, Mp, Np, Mp, 1.0, 0.0, true, true, false};
  g2.a_tri = +1;  // W'(i, kk) = W(kk, i) = 0 for kk < i
  return launch_gemm(g2, h->s, &h->launches);
}

static int sp_inverse(gpc_sparse* h, double* W, double* Out) {
  Dense d = sp_dense(h, W);
  return inverse_from_W(d, nullptr, h->Mp, h->Mp, Out, h->Mp, false);
}

#define SP_LAUNCH_CHECK(name)                                 \
  do {                                                        \
    h->launches++;                                            \
    GPC_CUDA_CHECK(cudaGetLastError());                       \
    if (trace_sync(name, h->s) != GPC_OK) return GPC_ERR_CUDA; \
  } while (0)

extern "C" {

int gpc_sparse_destroy(gpc_sparse* h) {
  if (!h) return GPC_OK;
  cudaSetDevice(h->device);
  if (h->s) cudaStreamSynchronize(h->s);
  double* bufs[] = {h->X, h->Mt, h->Xu, h->Kuu, h->Lu, h->Wu, h->Kuuinv, h->A, h->LA, h->WA, h->Ainv, h->gKuu, h->Kuf, h->V,
                    h->B, h->KufL, h->BS, h->gKuf, h->kdiag, h->q, h->lam, h->linv, h->gd, h->h, h->cd, h->a, h->ta, h->t1,
                    h->t2, h->Ba, h->tmpL, h->Tpool, h->symm_part, h->partial, h->gXu, h->scal, h->Xs, h->Ksu, h->Tq,
                    h->kss, h->pmu};
  for (double* b : bufs) cudaFree(b);
  cudaFree(h->info);
  if (h->hbuf) cudaFreeHost(h->hbuf);
  if (h->s) cudaStreamDestroy(h->s);
  delete h;
  return GPC_OK;
}

int gpc_sparse_create(gpc_sparse** out, int device, int approx, int64_t N, int M, int D, int dout) {
  if (!out || N < 1 || M < 1 || D < 1 || dout < 1 || (approx != SP_DTC && approx != SP_FITC && approx != SP_DTCVAR)) {
    set_error("gpc_sparse_create: bad arguments (approx: 1 DTC, 2 FITC, 4 DTCVAR as CGp.h:13-19)");
    return GPC_ERR_ARG;
  }
  *out = nullptr;
  int ndev = 0;
  GPC_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) {
    set_error("gpc_sparse_create: no such CUDA device");
    return GPC_ERR_CUDA;
  }
  GPC_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  GPC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("gpc_b200 is built for sm_100a only");
    return GPC_ERR_CUDA;
  }
  gpc_sparse* h = new gpc_sparse();
  memset(h, 0, sizeof(*h));
  h->device = device;
  h->approx = approx;
  h->N = N;
  h->Np = round_up(N, TILE);
  h->M = M;
  h->Mp = (int)round_up(M, TILE);
  h->D = D;
  h->d = dout;
  h->max_ctas = prop.multiProcessorCount * 2;
  const size_t Mp = (size_t)h->Mp, Np = (size_t)h->Np;
  auto alloc = [&](double** p, size_t n) -> bool {
    if (cudaMalloc(p, n * sizeof(double)) != cudaSuccess) return false;
    return cudaMemset(*p, 0, n * sizeof(double)) == cudaSuccess;
  };
  bool ok = cudaStreamCreateWithFlags(&h->s, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && alloc(&h->X, Np * D) && alloc(&h->Mt, Np * dout) && alloc(&h->Xu, Mp * D);
  double** mm[] = {&h->Kuu, &h->Lu, &h->Wu, &h->Kuuinv, &h->A, &h->LA, &h->WA, &h->Ainv, &h->gKuu};
  for (double** p : mm) ok = ok && alloc(p, Mp * Mp);
  double** mn[] = {&h->Kuf, &h->V, &h->B, &h->KufL, &h->BS, &h->gKuf};
  for (double** p : mn) ok = ok && alloc(p, Mp * Np);
  double** nv[] = {&h->kdiag, &h->q, &h->lam, &h->linv, &h->gd, &h->h, &h->cd};
  for (double** p : nv) ok = ok && alloc(p, Np);
  ok = ok && alloc(&h->a, Np * dout) && alloc(&h->ta, Np * dout);
  ok = ok && alloc(&h->t1, Mp * dout) && alloc(&h->t2, Mp * dout) && alloc(&h->Ba, Mp * dout);
  ok = ok && alloc(&h->tmpL, (Mp / 2 + TILE) * (Mp / 2 + TILE)) && alloc(&h->Tpool, potrf_inv_tspace((int64_t)Mp) + 16);
  ok = ok && alloc(&h->symm_part, (size_t)symm_chunks((int64_t)Mp) * 4 * Mp);
  ok = ok && alloc(&h->partial, (size_t)h->max_ctas * GPC_MAX_PARAMS) && alloc(&h->gXu, Mp * D);
  ok = ok && alloc(&h->scal, SS_G0 + 3 * GPC_MAX_PARAMS);
  ok = ok && cudaMalloc(&h->info, sizeof(int)) == cudaSuccess;
  ok = ok && cudaMallocHost(&h->hbuf, (4 * Np + SS_G0 + 3 * GPC_MAX_PARAMS + Mp * D) * sizeof(double)) == cudaSuccess;
  if (!ok) {
    set_error(std::string("gpc_sparse_create: ") + cudaGetErrorString(cudaGetLastError()));
    gpc_sparse_destroy(h);
    cudaGetLastError();
    return GPC_ERR_NOMEM;
  }
  *out = h;
  return GPC_OK;
}

int gpc_sparse_set_data(gpc_sparse* h, const double* X, int64_t ldx, const double* M, int64_t ldm) {
  if (!h || !X || !M || ldx < h->N || ldm < h->N) {
    set_error("gpc_sparse_set_data: bad arguments");
    return GPC_ERR_ARG;
  }
  GPC_CUDA_CHECK(cudaSetDevice(h->device));
  GPC_CUDA_CHECK(cudaMemsetAsync(h->X, 0, (size_t)h->Np * h->D * sizeof(double), h->s));
  GPC_CUDA_CHECK(cudaMemsetAsync(h->Mt, 0, (size_t)h->Np * h->d * sizeof(double), h->s));
  GPC_CUDA_CHECK(cudaMemcpy2DAsync(h->X, h->Np * sizeof(double), X, ldx * sizeof(double), h->N * sizeof(double), h->D,
                                   cudaMemcpyHostToDevice, h->s));
  GPC_CUDA_CHECK(cudaMemcpy2DAsync(h->Mt, h->Np * sizeof(double), M, ldm * sizeof(double), h->N * sizeof(double), h->d,
                                   cudaMemcpyHostToDevice, h->s));
  GPC_CUDA_CHECK(cudaStreamSynchronize(h->s));
  h->haveData = true;
  h->haveEval = false;
  return GPC_OK;
}

int gpc_sparse_eval(gpc_sparse* h, const gpc_kcomp* comps, int ncomp, const double* Xu, int64_t ldxu, double beta,
                    double* out, double* gparams, double* gXu, double* gbeta_out) {
  if (!h || !Xu || ldxu < h->M || !(beta > 0.0)) {
    set_error("gpc_sparse_eval: bad arguments");
    return GPC_ERR_ARG;
  }
  if (!h->haveData) {
    set_error("gpc_sparse_eval: X and m have not been set");
    return GPC_ERR_STATE;
  }
  GPC_CUDA_CHECK(cudaSetDevice(h->device));
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, h->D, &ks));
  const int64_t Mp = h->Mp, Np = h->Np, N = h->N;
  const int M = h->M, d = h->d, D = h->D;
  const bool fitc = h->approx == SP_FITC, dtcvar = h->approx == SP_DTCVAR;
  cudaStream_t s = h->s;
  h->haveEval = false;
  h->beta = beta;
  // ---- 1. X_u, K_uu (diagonal through diagComputeElement: white included, CGp.cpp:721), K_uf (computeElement), k_ii
  GPC_CUDA_CHECK(cudaMemsetAsync(h->Xu, 0, (size_t)Mp * D * sizeof(double), s));
  GPC_CUDA_CHECK(cudaMemcpy2DAsync(h->Xu, Mp * sizeof(double), Xu, ldxu * sizeof(double), M * sizeof(double), D,
                                   cudaMemcpyHostToDevice, s));
  GPC_CHECK(launch_kbuild(ks, h->Xu, Mp, M, Mp, h->Kuu, Mp, s, &h->launches));
  GPC_CHECK(launch_kcross(ks, h->Xu, Mp, M, Mp, h->X, Np, N, Np, h->Kuf, Mp, s, &h->launches));
  GPC_CUDA_CHECK(cudaMemsetAsync(h->kdiag, 0, (size_t)Np * sizeof(double), s));
  if (fitc || dtcvar) GPC_CHECK(launch_kdiag(ks, h->X, Np, N, h->kdiag, s, &h->launches));
  // ---- 2. K_uu = L_u L_u', W_u; B = K_uu^-1 K_uf; q_ii
  double logdetKuu = 0.0, logdetA = 0.0, jitKuu = 0.0, jitA = 0.0;
  {
    int rc = sp_factor(h, h->Kuu, h->Lu, h->Wu, &logdetKuu, &jitKuu);
    if (rc != GPC_OK) return rc;
  }
  GPC_CHECK(sp_solve(h, h->Wu, h->Kuf, h->V, h->B));
  const unsigned cgrid = (unsigned)((Np + 7) / 8);
  sp_coldot_kernel<<<cgrid, 256, 0, s>>>(h->q, h->Kuf, h->B, Mp, Np);
  SP_LAUNCH_CHECK("sp_coldot_kernel");
  // ---- 3. Lambda, K_uf Lambda^-1
  sp_lambda_kernel<<<(unsigned)((Np + 255) / 256), 256, 0, s>>>(h->lam, h->linv, h->kdiag, h->q, beta, fitc ? 1 : 0, N, Np);
  SP_LAUNCH_CHECK("sp_lambda_kernel");
  dim3 gmn((unsigned)Np, (unsigned)((Mp + 255) / 256));
  sp_colscale_kernel<<<gmn, 256, 0, s>>>(h->KufL, h->Kuf, Mp, Np, h->linv, 0.0);
  SP_LAUNCH_CHECK("sp_colscale_kernel");
  // ---- 4. A = K_uu + K_uf Lambda^-1 K_fu (lower), its factor and inverse
  GPC_CHECK(launch_copy_lower(h->Kuu, Mp, h->A, Mp, Mp, s, &h->launches));
  {
    GemmCall g{h->KufL, h->Kuf, h->A, Mp, Mp, Mp, Mp, Mp, Np, 1.0, 1.0, false, false, true};
    GPC_CHECK(launch_gemm(g, s, &h->launches));
  }
  {
    int rc = sp_factor(h, h->A, h->LA, h->WA, &logdetA, &jitA);
    if (rc != GPC_OK) return rc;
  }
  GPC_CHECK(sp_inverse(h, h->WA, h->Ainv));
  GPC_CHECK(sp_inverse(h, h->Wu, h->Kuuinv));
  // ---- 5. BS = A^-1 K_uf Lambda^-1; a = Sigma^-1 m; Ba = BS m
  GPC_CHECK(sp_solve(h, h->WA, h->KufL, h->V, h->BS));
  GPC_CUDA_CHECK(cudaMemsetAsync(h->t1, 0, (size_t)Mp * d * sizeof(double), s));
  GPC_CUDA_CHECK(cudaMemsetAsync(h->Ba, 0, (size_t)Mp * d * sizeof(double), s));
  {
    const int64_t chunk = 512;
    dim3 g((unsigned)((Mp + 127) / 128), (unsigned)((Np + chunk - 1) / chunk));
    sp_matvec_rows_kernel<<<g, 128, 0, s>>>(h->KufL, Mp, Np, h->Mt, Np, d, h->t1, Mp, chunk);
    SP_LAUNCH_CHECK("sp_matvec_rows_kernel");
  }
  // t2 = A^-1 t1 as W_A' (W_A t1), NOT through the explicit A^-1: t1 = K_uf Lambda^-1 m lives in A's large eigen-directions, so
  // A^-1 t1 is tiny against |A^-1| |t1| and the product with the full inverse would lose cond(A) eps of the quadratic form
  // (the reference's own dsymv_ with Ainv, CGp.cpp:948-949, does; the triangular products do not)
  GPC_CHECK(launch_gemv_rows(h->WA, Mp, Mp, Mp, h->t1, Mp, d, h->Ba, Mp, s, &h->launches));  // Ba as scratch: u = W_A t1
  sp_matvec_cols_kernel<<<(unsigned)((Mp + 7) / 8), 256, 0, s>>>(h->WA, Mp, Mp, h->Ba, Mp, d, h->t2, Mp);
  SP_LAUNCH_CHECK("sp_matvec_cols_kernel");
  GPC_CUDA_CHECK(cudaMemsetAsync(h->Ba, 0, (size_t)Mp * d * sizeof(double), s));
  {
    const int64_t chunk = 512;
    dim3 g((unsigned)((Mp + 127) / 128), (unsigned)((Np + chunk - 1) / chunk));
    sp_matvec_rows_kernel<<<g, 128, 0, s>>>(h->BS, Mp, Np, h->Mt, Np, d, h->Ba, Mp, chunk);
    SP_LAUNCH_CHECK("sp_matvec_rows_kernel");
  }
  sp_matvec_cols_kernel<<<cgrid, 256, 0, s>>>(h->KufL, Mp, Np, h->t2, Mp, d, h->ta, Np);
  SP_LAUNCH_CHECK("sp_matvec_cols_kernel");
  sp_a_kernel<<<(unsigned)((Np + 255) / 256), 256, 0, s>>>(h->a, h->Mt, h->ta, h->linv, Np, d);
  SP_LAUNCH_CHECK("sp_a_kernel");
  GPC_CUDA_CHECK(cudaMemsetAsync(h->scal + SS_QUAD, 0, sizeof(double), s));
  GPC_CHECK(launch_dot(h->a, h->Mt, Np * d, h->scal + SS_QUAD, s, &h->launches));
  // ---- 6. diag(G), h, dL/dK_uf, dL/dK_uu
  const double c = dtcvar ? -0.5 * (double)d * beta : 0.0;
  sp_coldot_kernel<<<cgrid, 256, 0, s>>>(h->cd, h->KufL, h->BS, Mp, Np);
  SP_LAUNCH_CHECK("sp_coldot_kernel");
  sp_gd_kernel<<<(unsigned)((Np + 255) / 256), 256, 0, s>>>(h->gd, h->h, h->linv, h->cd, h->a, N, Np, d, fitc ? 1 : 0, c);
  SP_LAUNCH_CHECK("sp_gd_kernel");
  sp_gkuf_kernel<<<gmn, 256, 0, s>>>(h->gKuf, h->BS, h->B, h->Ba, h->a, h->gd, Mp, Np, d, fitc ? 1 : 0, c);
  SP_LAUNCH_CHECK("sp_gkuf_kernel");
  dim3 gmm((unsigned)Mp, (unsigned)((Mp + 255) / 256));
  sp_gkuu_kernel<<<gmm, 256, 0, s>>>(h->gKuu, h->Kuuinv, h->Ainv, h->Ba, Mp, d);
  SP_LAUNCH_CHECK("sp_gkuu_kernel");
  if (fitc || dtcvar) {  // + B Z B',  Z = diag(gd) (FITC) or c I (DTCVAR)
    if (fitc) {
      sp_colscale_kernel<<<gmn, 256, 0, s>>>(h->V, h->B, Mp, Np, h->gd, 0.0);
    } else {  // constant weight: w = 0 + c
      GPC_CUDA_CHECK(cudaMemsetAsync(h->ta, 0, (size_t)Np * sizeof(double), s));
      sp_colscale_kernel<<<gmn, 256, 0, s>>>(h->V, h->B, Mp, Np, h->ta, c);
    }
    SP_LAUNCH_CHECK("sp_colscale_kernel");
    GemmCall g{h->V, h->B, h->gKuu, Mp, Mp, Mp, Mp, Mp, Np, 1.0, 1.0, false, false, true};
    GPC_CHECK(launch_gemm(g, s, &h->launches));
  }
  // ---- 7. kernel-parameter gradients (natural) and dL/dX_u: K_uu part (symmetric, white on its diagonal), K_uf part (cross)
  //         and the diagonal term sum_i h_i dk_ii/dtheta
  double* g_uu = h->scal + SS_G0;
  double* g_uf = g_uu + GPC_MAX_PARAMS;
  double* g_dg = g_uf + GPC_MAX_PARAMS;
  GPC_CUDA_CHECK(cudaMemsetAsync(g_uu, 0, 3 * GPC_MAX_PARAMS * sizeof(double), s));
  GPC_CUDA_CHECK(cudaMemsetAsync(h->gXu, 0, (size_t)Mp * D * sizeof(double), s));
  GPC_CHECK(launch_grad(ks, h->Xu, Mp, M, Mp, h->gKuu, Mp, nullptr, 0, 0, 1, h->partial, h->max_ctas, g_uu, h->gXu, Mp, s,
                        &h->launches));
  {
    GradCross cx{1, h->X, Np, N};
    GPC_CHECK(launch_grad(ks, h->Xu, Mp, M, Mp, h->gKuf, Mp, nullptr, 0, 0, 1, h->partial, h->max_ctas, g_uf, h->gXu, Mp, s,
                          &h->launches, -1, 0, nullptr, &cx));
  }
  if (fitc || dtcvar) {
    sp_kdiag_grad_kernel<<<h->max_ctas, 256, 0, s>>>(ks, h->X, Np, N, h->h, g_dg);
    SP_LAUNCH_CHECK("sp_kdiag_grad_kernel");
  }
  // ---- 8. results: scalars, the three gradient vectors, dL/dX_u, and the N-vectors reduced on the host
  double* hb = h->hbuf;
  double* h_lam = hb;
  double* h_gd = hb + Np;
  double* h_kd = hb + 2 * Np;
  double* h_q = hb + 3 * Np;
  double* h_sc = hb + 4 * Np;
  double* h_gx = h_sc + SS_G0 + 3 * GPC_MAX_PARAMS;
  GPC_CUDA_CHECK(cudaMemcpyAsync(h_lam, h->lam, N * sizeof(double), cudaMemcpyDeviceToHost, s));
  GPC_CUDA_CHECK(cudaMemcpyAsync(h_gd, h->gd, N * sizeof(double), cudaMemcpyDeviceToHost, s));
  GPC_CUDA_CHECK(cudaMemcpyAsync(h_kd, h->kdiag, N * sizeof(double), cudaMemcpyDeviceToHost, s));
  GPC_CUDA_CHECK(cudaMemcpyAsync(h_q, h->q, N * sizeof(double), cudaMemcpyDeviceToHost, s));
  GPC_CUDA_CHECK(cudaMemcpyAsync(h_sc, h->scal, (SS_G0 + 3 * GPC_MAX_PARAMS) * sizeof(double), cudaMemcpyDeviceToHost, s));
  GPC_CUDA_CHECK(cudaMemcpyAsync(h_gx, h->gXu, (size_t)Mp * D * sizeof(double), cudaMemcpyDeviceToHost, s));
  GPC_CUDA_CHECK(cudaStreamSynchronize(s));
  long double sl = 0.0L, sg = 0.0L, skq = 0.0L;
  for (int64_t i = 0; i < N; i++) {
    sl += logl((long double)h_lam[i]);
    sg += h_gd[i];
    skq += (long double)h_kd[i] - (long double)h_q[i];
  }
  const double logdet = (double)sl - logdetKuu + logdetA;
  const double quad = h_sc[SS_QUAD];
  const double HALFLOG2PI = 0.91893853320467274178;
  double ll = -0.5 * ((double)d * logdet + quad) - (double)d * (double)N * HALFLOG2PI;
  if (fitc) ll -= (double)d * (double)N * HALFLOG2PI;  // the reference counts the constant twice (CGp.cpp:963 and :1012)
  double gbeta = -(double)sg / (beta * beta);
  if (dtcvar) {
    ll -= 0.5 * (double)d * beta * (double)skq;  // CGp.cpp:955-956 with diagD of :788-791
    gbeta += -0.5 * (double)d * (double)skq;
  }
  if (out) {
    out[0] = ll;
    out[1] = logdet;
    out[2] = quad;
    out[3] = (double)skq;
    out[4] = jitKuu;
    out[5] = jitA;
  }
  if (gparams)
    for (int i = 0; i < ks.nparams; i++)
      gparams[i] = h_sc[SS_G0 + i] + h_sc[SS_G0 + GPC_MAX_PARAMS + i] + h_sc[SS_G0 + 2 * GPC_MAX_PARAMS + i];
  if (gXu)
    for (int k = 0; k < D; k++)
      for (int i = 0; i < M; i++) gXu[i + (size_t)k * M] = h_gx[i + (size_t)k * Mp];
  if (gbeta_out) *gbeta_out = gbeta;
  h->haveEval = true;
  return GPC_OK;
}

// posterior at Xs (Ns x D): mu (Ns x d, ld Ns) = K_*u A^-1 K_uf Lambda^-1 m, var (Ns) = k_** - k_*u' (K_uu^-1 - A^-1) k_*u + 1/beta
// in the space of m (CGp::updateAlpha CGp.cpp:490-521, _posteriorVar :584-599); needs a gpc_sparse_eval at the current parameters
int gpc_sparse_posterior(gpc_sparse* h, const gpc_kcomp* comps, int ncomp, const double* Xs, int64_t Ns, int64_t ldxs,
                         double* mu, double* var) {
  if (!h || !Xs || !mu || Ns < 1 || ldxs < Ns) {
    set_error("gpc_sparse_posterior: bad arguments");
    return GPC_ERR_ARG;
  }
  if (!h->haveEval) {
    set_error("gpc_sparse_posterior: needs a successful gpc_sparse_eval first");
    return GPC_ERR_STATE;
  }
  GPC_CUDA_CHECK(cudaSetDevice(h->device));
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, h->D, &ks));
  const int64_t Nsp = round_up(Ns, TILE), Mp = h->Mp;
  cudaStream_t s = h->s;
  if (h->ps_cap < Nsp) {
    cudaFree(h->Xs); cudaFree(h->Ksu); cudaFree(h->Tq); cudaFree(h->kss); cudaFree(h->pmu);
    h->Xs = h->Ksu = h->Tq = h->kss = h->pmu = nullptr;
    h->ps_cap = 0;
    GPC_CUDA_CHECK(cudaMalloc(&h->Xs, (size_t)Nsp * h->D * sizeof(double)));
    GPC_CUDA_CHECK(cudaMalloc(&h->Ksu, (size_t)Nsp * Mp * sizeof(double)));
    GPC_CUDA_CHECK(cudaMalloc(&h->Tq, (size_t)Nsp * Mp * sizeof(double)));
    GPC_CUDA_CHECK(cudaMalloc(&h->kss, (size_t)Nsp * sizeof(double)));
    GPC_CUDA_CHECK(cudaMalloc(&h->pmu, (size_t)Nsp * h->d * sizeof(double)));
    h->ps_cap = Nsp;
  }
  GPC_CUDA_CHECK(cudaMemsetAsync(h->Xs, 0, (size_t)Nsp * h->D * sizeof(double), s));
  GPC_CUDA_CHECK(cudaMemcpy2DAsync(h->Xs, Nsp * sizeof(double), Xs, ldxs * sizeof(double), Ns * sizeof(double), h->D,
                                   cudaMemcpyHostToDevice, s));
  GPC_CHECK(launch_kcross(ks, h->Xs, Nsp, Ns, Nsp, h->Xu, Mp, h->M, Mp, h->Ksu, Nsp, s, &h->launches));
  GPC_CHECK(launch_gemv_rows(h->Ksu, Nsp, Ns, h->M, h->t2, Mp, h->d, h->pmu, Nsp, s, &h->launches));
  GPC_CUDA_CHECK(cudaMemcpy2DAsync(mu, Ns * sizeof(double), h->pmu, Nsp * sizeof(double), Ns * sizeof(double), h->d,
                                   cudaMemcpyDeviceToHost, s));
  if (var) {
    sp_sub_kernel<<<1024, 256, 0, s>>>(h->gKuu, h->Kuuinv, h->Ainv, (int64_t)Mp * Mp);  // E = K_uu^-1 - A^-1 (gKuu as scratch)
    SP_LAUNCH_CHECK("sp_sub_kernel");
    GemmCall g{h->Ksu, h->gKuu, h->Tq, Nsp, Mp, Nsp, Nsp, Mp, Mp, 1.0, 0.0, false, true, false};
    GPC_CHECK(launch_gemm(g, s, &h->launches));
    GPC_CHECK(launch_kdiag(ks, h->Xs, Nsp, Ns, h->kss, s, &h->launches));
    sp_var_kernel<<<(unsigned)((Ns + 127) / 128), 128, 0, s>>>(h->kss, h->kss, h->Tq, h->Ksu, Ns, Nsp, h->M, 1.0 / h->beta);
    SP_LAUNCH_CHECK("sp_var_kernel");
    GPC_CUDA_CHECK(cudaMemcpyAsync(var, h->kss, Ns * sizeof(double), cudaMemcpyDeviceToHost, s));
  }
  GPC_CUDA_CHECK(cudaStreamSynchronize(s));
  return GPC_OK;
}

int64_t gpc_sparse_launch_count(gpc_sparse* h) { return h ? h->launches : 0; }

}  // extern "C"
