// lapack_api.cu -- CMatrix-level drop-ins for the lapack.h calls on the hot path (host pointers, LAPACK
// argument meaning), each staged through TILE-padded device buffers and executed by the same DMMA engine,
// plus the measurement helpers bench.py uses.  See include/gpc_b200.h for the reference lines replaced.
#include <math.h>
#include <string.h>
#include <vector>
#include "common.cuh"

using namespace gpc;

#define GPC_CHECK(expr)            \
  do {                             \
    int _rc = (expr);              \
    if (_rc != GPC_OK) return _rc; \
  } while (0)

namespace {

// Device scratch of one CMatrix-level call.  The stream and the buffers come from a per-thread cache (per device):
// a level-0 / level-1 caller makes hundreds of small calls per optimiser iteration, and a
// cudaStreamCreate + cudaMalloc/cudaFree round per call cost more than the work.  Buffers are handed out best-fit,
// returned when the call ends, and the cache is dropped when it holds more than 4 GB of idle memory.
struct ScratchCache {
  struct Buf {
    void* p;
    size_t bytes;
    bool busy;
  };
  cudaStream_t s = nullptr;
  std::vector<Buf> bufs;
  ~ScratchCache() {}  // process exit: the driver reclaims everything; cudaFree here could run after the context died
};
static thread_local ScratchCache g_scratch[64];

struct Scratch {  // RAII view of the cache for one call
  cudaStream_t s = nullptr;
  ScratchCache* cache = nullptr;
  int64_t launches = 0;
  int init(int device) {
    int ndev = 0;
    GPC_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev || device >= 64) {
      set_error("no such CUDA device");
      return GPC_ERR_CUDA;
    }
    GPC_CUDA_CHECK(cudaSetDevice(device));
    cache = &g_scratch[device];
    if (!cache->s) GPC_CUDA_CHECK(cudaStreamCreateWithFlags(&cache->s, cudaStreamNonBlocking));
    s = cache->s;
    return GPC_OK;
  }
  int alloc(double** p, size_t elems, bool zero) {
    const size_t need = (elems ? elems : 1) * sizeof(double);
    int best = -1;
    for (size_t i = 0; i < cache->bufs.size(); i++) {
      const ScratchCache::Buf& b = cache->bufs[i];
      if (!b.busy && b.bytes >= need && (best < 0 || b.bytes < cache->bufs[(size_t)best].bytes)) best = (int)i;
    }
    if (best >= 0 && cache->bufs[(size_t)best].bytes <= 4 * need + (1 << 20)) {
      cache->bufs[(size_t)best].busy = true;
      *p = (double*)cache->bufs[(size_t)best].p;
    } else {
      void* q = nullptr;
      GPC_CUDA_CHECK(cudaMalloc(&q, need));
      cache->bufs.push_back({q, need, true});
      *p = (double*)q;
    }
    if (zero) GPC_CUDA_CHECK(cudaMemsetAsync(*p, 0, elems * sizeof(double), s));
    return GPC_OK;
  }
  ~Scratch() {
    if (!cache) return;
    if (s) cudaStreamSynchronize(s);
    size_t idle = 0;
    for (auto& b : cache->bufs) {
      b.busy = false;
      idle += b.bytes;
    }
    if (idle > ((size_t)4 << 30)) {
      for (auto& b : cache->bufs) cudaFree(b.p);
      cache->bufs.clear();
    }
  }
};

inline bool is(char c, char u) { return c == u || c == (char)(u + 32); }

int up(Scratch& sc, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return GPC_OK;
  GPC_CUDA_CHECK(cudaMemcpy2DAsync(dst, ldd * sizeof(double), src, lds * sizeof(double), rows * sizeof(double), cols,
                                   cudaMemcpyHostToDevice, sc.s));
  return GPC_OK;
}
int down(Scratch& sc, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return GPC_OK;
  GPC_CUDA_CHECK(cudaMemcpy2DAsync(dst, ldd * sizeof(double), src, lds * sizeof(double), rows * sizeof(double), cols,
                                   cudaMemcpyDeviceToHost, sc.s));
  GPC_CUDA_CHECK(cudaStreamSynchronize(sc.s));
  return GPC_OK;
}

// Stage a host triangular / symmetric n x n matrix so that the referenced triangle sits in the LOWER part of a
// TILE-padded device matrix (identity in the padding).  uplo 'U' data is transposed on the device.
int stage_lower(Scratch& sc, char uplo, int64_t n, const double* A, int64_t lda, double** out, int64_t* npad) {
  int64_t np = round_up(n, TILE);
  double* d0;
  GPC_CHECK(sc.alloc(&d0, (size_t)np * np, true));
  GPC_CHECK(up(sc, d0, np, A, lda, n, n));
  double* dl = d0;
  if (is(uplo, 'U')) {
    GPC_CHECK(sc.alloc(&dl, (size_t)np * np, true));
    GPC_CHECK(launch_transpose(d0, np, dl, np, n, n, sc.s, &sc.launches));
  }
  GPC_CHECK(launch_set_identity_pad(dl, np, n, np, sc.s, &sc.launches));
  *out = dl;
  *npad = np;
  return GPC_OK;
}

}  // namespace

extern "C" {

int gpc_dpotrf(int device, char uplo, int64_t n, double* A, int64_t lda, int* info) {
  if (!A || n < 1 || lda < n || !(is(uplo, 'U') || is(uplo, 'L'))) {
    set_error("gpc_dpotrf: bad arguments");
    return GPC_ERR_ARG;
  }
  Scratch sc;
  GPC_CHECK(sc.init(device));
  double* dl;
  int64_t np;
  GPC_CHECK(stage_lower(sc, uplo, n, A, lda, &dl, &np));
  Dense d;
  d.s = sc.s;
  d.launches = &sc.launches;
  double* scal;
  GPC_CHECK(sc.alloc(&d.Dinv, (size_t)np * TILE, false));
  GPC_CHECK(sc.alloc(&scal, 2, true));
  d.info = (int*)(scal + 1);
  d.logdet = scal;
  d.W = nullptr;
  d.nvalid = n;
  GPC_CHECK(potrf_rec(d, dl, np, np, 0));
  int hinfo = 0;
  GPC_CUDA_CHECK(cudaMemcpyAsync(&hinfo, d.info, sizeof(int), cudaMemcpyDeviceToHost, sc.s));
  GPC_CUDA_CHECK(cudaStreamSynchronize(sc.s));
  if (info) *info = hinfo;
  // write back only the `uplo` triangle (LAPACK leaves the other one untouched; CMatrix::chol zeroes it itself)
  std::vector<double> h((size_t)n * n);
  GPC_CHECK(down(sc, h.data(), n, dl, np, n, n));
  if (is(uplo, 'L')) {
    for (int64_t j = 0; j < n; j++)
      for (int64_t i = j; i < n; i++) A[i + j * lda] = h[i + j * n];
  } else {
    for (int64_t j = 0; j < n; j++)
      for (int64_t i = 0; i <= j; i++) A[i + j * lda] = h[j + i * n];
  }
  return hinfo;
}

int gpc_dpotri(int device, char uplo, int64_t n, double* A, int64_t lda, int* info) {
  if (!A || n < 1 || lda < n || !(is(uplo, 'U') || is(uplo, 'L'))) {
    set_error("gpc_dpotri: bad arguments");
    return GPC_ERR_ARG;
  }
  for (int64_t i = 0; i < n; i++)
    if (A[i + i * lda] == 0.0) {  // LAPACK: info = i if the (i,i) element of the factor is zero
      if (info) *info = (int)(i + 1);
      return (int)(i + 1);
    }
  Scratch sc;
  GPC_CHECK(sc.init(device));
  double* dl;
  int64_t np;
  GPC_CHECK(stage_lower(sc, uplo, n, A, lda, &dl, &np));
  GPC_CHECK(launch_zero_upper(dl, np, np, sc.s, &sc.launches));
  Dense d;
  d.s = sc.s;
  d.launches = &sc.launches;
  d.info = nullptr;
  d.logdet = nullptr;
  d.nvalid = n;
  GPC_CHECK(sc.alloc(&d.Dinv, (size_t)np * TILE, false));
  GPC_CHECK(sc.alloc(&d.W, potri_workspace(np) + 16, false));
  for (int64_t b = 0; b < np / TILE; b++)
    GPC_CHECK(launch_trtri_leaf(dl + b * TILE + b * TILE * np, np, d.Dinv + b * TILE * TILE, sc.s, &sc.launches));
  double* inv;
  GPC_CHECK(sc.alloc(&inv, (size_t)np * np, false));
  GPC_CHECK(potri_rec(d, dl, np, np, inv, np, 0));
  std::vector<double> hv((size_t)n * n);
  GPC_CHECK(down(sc, hv.data(), n, inv, np, n, n));
  if (is(uplo, 'L')) {
    for (int64_t j = 0; j < n; j++)
      for (int64_t i = j; i < n; i++) A[i + j * lda] = hv[i + j * n];
  } else {
    for (int64_t j = 0; j < n; j++)
      for (int64_t i = 0; i <= j; i++) A[i + j * lda] = hv[i + j * n];
  }
  if (info) *info = 0;
  return GPC_OK;
}

int gpc_dtrsm(int device, char side, char uplo, char transa, char diag, int64_t m, int64_t n, double alpha,
              const double* A, int64_t lda, double* B, int64_t ldb) {
  bool left = is(side, 'L'), upper = is(uplo, 'U'), tr = is(transa, 'T') || is(transa, 'C'), unit = is(diag, 'U');
  int64_t k = left ? m : n;
  if (!A || !B || m < 1 || n < 1 || lda < k || ldb < m) {
    set_error("gpc_dtrsm: bad arguments");
    return GPC_ERR_ARG;
  }
  Scratch sc;
  GPC_CHECK(sc.init(device));
  // Lo = lower-triangular staging of A (A itself, or A' when A is upper)
  double* dl;
  int64_t kp;
  if (unit) {
    std::vector<double> Au((size_t)k * k);
    for (int64_t j = 0; j < k; j++)
      for (int64_t i = 0; i < k; i++) Au[i + j * k] = (i == j) ? 1.0 : A[i + j * lda];
    GPC_CHECK(stage_lower(sc, uplo, k, Au.data(), k, &dl, &kp));
    GPC_CUDA_CHECK(cudaStreamSynchronize(sc.s));
  } else {
    GPC_CHECK(stage_lower(sc, uplo, k, A, lda, &dl, &kp));
  }
  GPC_CHECK(launch_zero_upper(dl, kp, kp, sc.s, &sc.launches));
  Dense d;
  d.s = sc.s;
  d.launches = &sc.launches;
  d.info = nullptr;
  d.logdet = nullptr;
  d.W = nullptr;
  d.nvalid = k;
  GPC_CHECK(sc.alloc(&d.Dinv, (size_t)kp * TILE, false));
  for (int64_t b = 0; b < kp / TILE; b++)
    GPC_CHECK(launch_trtri_leaf(dl + b * TILE + b * TILE * kp, kp, d.Dinv + b * TILE * TILE, sc.s, &sc.launches));
  // right-hand side staged as R (rows x kp) with the triangular dimension along the columns:
  //   side R: R = alpha*B (m x n);  side L: R = alpha*B' (n x m)
  int64_t rows = left ? n : m;
  int64_t rp = round_up(rows, TILE);
  double *d0, *R;
  int64_t mp = round_up(m, TILE), np_ = round_up(n, TILE);
  GPC_CHECK(sc.alloc(&d0, (size_t)mp * np_, true));
  GPC_CHECK(up(sc, d0, mp, B, ldb, m, n));
  GPC_CHECK(sc.alloc(&R, (size_t)rp * kp, true));
  if (left)
    GPC_CHECK(launch_transpose(d0, mp, R, rp, m, n, sc.s, &sc.launches));
  else
    GPC_CHECK(launch_copy_block(d0, mp, R, rp, m, n, 1.0, sc.s, &sc.launches));
  // which canonical right-sided solve: X Lo' = R (rlt) or X Lo = R (rln)
  bool use_rlt = left ? (upper ? tr : !tr) : (upper ? !tr : tr);
  if (use_rlt)
    GPC_CHECK(trsm_rlt(d, R, rp, rp, dl, kp, kp, 0));
  else
    GPC_CHECK(trsm_rln(d, R, rp, rp, dl, kp, kp, 0));
  std::vector<double> h((size_t)rows * k);
  GPC_CHECK(down(sc, h.data(), rows, R, rp, rows, k));
  if (left) {
    for (int64_t j = 0; j < n; j++)
      for (int64_t i = 0; i < m; i++) B[i + j * ldb] = alpha * h[j + i * rows];
  } else {
    for (int64_t j = 0; j < n; j++)
      for (int64_t i = 0; i < m; i++) B[i + j * ldb] = alpha * h[i + j * rows];
  }
  return GPC_OK;
}

// P = op(A) op(B) on the device into a host buffer (m x n, ld m)
static int device_product(int device, bool ta, bool tb, int64_t m, int64_t n, int64_t k, const double* A, int64_t lda,
                          const double* B, int64_t ldb, std::vector<double>& P) {
  Scratch sc;
  GPC_CHECK(sc.init(device));
  int64_t mp = round_up(m, TILE), np = round_up(n, TILE), kp = round_up(k, 16);
  double *dA, *dB, *dC;
  // A as stored: (m x k) if !ta else (k x m)
  int64_t ar = ta ? kp : mp, ac = ta ? mp : kp;
  int64_t br = tb ? np : kp, bc = tb ? kp : np;
  GPC_CHECK(sc.alloc(&dA, (size_t)ar * ac, true));
  GPC_CHECK(sc.alloc(&dB, (size_t)br * bc, true));
  GPC_CHECK(sc.alloc(&dC, (size_t)mp * np, false));
  GPC_CHECK(up(sc, dA, ar, A, lda, ta ? k : m, ta ? m : k));
  GPC_CHECK(up(sc, dB, br, B, ldb, tb ? n : k, tb ? k : n));
  GemmCall g{dA, dB, dC, ar, br, mp, mp, np, kp, 1.0, 0.0, ta, !tb, false};
  GPC_CHECK(launch_gemm(g, sc.s, &sc.launches));
  P.resize((size_t)m * n);
  return down(sc, P.data(), m, dC, mp, m, n);
}

int gpc_dgemm(int device, char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
              int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
  bool ta = is(transa, 'T') || is(transa, 'C'), tb = is(transb, 'T') || is(transb, 'C');
  if (!A || !B || !C || m < 1 || n < 1 || k < 1 || ldc < m) {
    set_error("gpc_dgemm: bad arguments");
    return GPC_ERR_ARG;
  }
  std::vector<double> P;
  GPC_CHECK(device_product(device, ta, tb, m, n, k, A, lda, B, ldb, P));
  for (int64_t j = 0; j < n; j++)
    for (int64_t i = 0; i < m; i++) {
      double c = (beta == 0.0) ? 0.0 : beta * C[i + j * ldc];
      C[i + j * ldc] = alpha * P[i + j * m] + c;
    }
  return GPC_OK;
}

int gpc_dsyrk(int device, char uplo, char trans, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
              double beta, double* C, int64_t ldc) {
  bool t = is(trans, 'T') || is(trans, 'C');
  if (!A || !C || n < 1 || k < 1 || ldc < n) {
    set_error("gpc_dsyrk: bad arguments");
    return GPC_ERR_ARG;
  }
  std::vector<double> P;
  // 'N': C = alpha A A' + beta C (A n x k);  'T': C = alpha A' A + beta C (A k x n)
  GPC_CHECK(device_product(device, t, !t, n, n, k, A, lda, A, lda, P));
  bool up_ = is(uplo, 'U');
  for (int64_t j = 0; j < n; j++) {
    int64_t i0 = up_ ? 0 : j, i1 = up_ ? j + 1 : n;
    for (int64_t i = i0; i < i1; i++) {
      double c = (beta == 0.0) ? 0.0 : beta * C[i + j * ldc];
      C[i + j * ldc] = alpha * P[i + j * n] + c;
    }
  }
  return GPC_OK;
}

int gpc_dsymv(int device, char uplo, int64_t n, double alpha, const double* A, int64_t lda, const double* x,
              double beta, double* y) {
  if (!A || !x || !y || n < 1 || lda < n) {
    set_error("gpc_dsymv: bad arguments");
    return GPC_ERR_ARG;
  }
  Scratch sc;
  GPC_CHECK(sc.init(device));
  double* dl;
  int64_t np;
  GPC_CHECK(stage_lower(sc, uplo, n, A, lda, &dl, &np));
  GPC_CHECK(launch_mirror_lower(dl, np, np, sc.s, &sc.launches));
  double *dx, *dy;
  GPC_CHECK(sc.alloc(&dx, (size_t)np, true));
  GPC_CHECK(sc.alloc(&dy, (size_t)np, true));
  GPC_CHECK(up(sc, dx, np, x, n, n, 1));
  double* part;
  GPC_CHECK(sc.alloc(&part, (size_t)symm_chunks(n) * 4 * np, false));
  GPC_CHECK(launch_symm_small(dl, np, dx, np, dy, np, n, 1, part, sc.s, &sc.launches));
  std::vector<double> h((size_t)n);
  GPC_CHECK(down(sc, h.data(), n, dy, np, n, 1));
  for (int64_t i = 0; i < n; i++) y[i] = alpha * h[i] + ((beta == 0.0) ? 0.0 : beta * y[i]);
  return GPC_OK;
}

// A := alpha x x' + A on the `uplo` triangle (dsyr_, lapack.h:154-160; CMatrix::syr CMatrix.h:526-533 -- the rank-one term
// of CGp::updateCovGradient, CGp.cpp:672-674).  HBM-bound: one pass over the triangle.
__global__ void syr_lower_kernel(double* __restrict__ A, int64_t lda, int64_t n, double alpha, const double* __restrict__ x) {
  const int64_t j = blockIdx.y;
  const double axj = alpha * x[j];
  for (int64_t i = j + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    A[i + j * lda] = fma(x[i], axj, A[i + j * lda]);
}
}  // extern "C"
extern "C" int gpc_dsyr(int device, char uplo, int64_t n, double alpha, const double* x, int64_t incx, double* A, int64_t lda) {
  if (!A || !x || n < 1 || lda < n || incx < 1 || !(is(uplo, 'U') || is(uplo, 'L'))) {
    set_error("gpc_dsyr: bad arguments");
    return GPC_ERR_ARG;
  }
  Scratch sc;
  GPC_CHECK(sc.init(device));
  double* dl;
  int64_t np;
  GPC_CHECK(stage_lower(sc, uplo, n, A, lda, &dl, &np));
  std::vector<double> hx((size_t)n);
  for (int64_t i = 0; i < n; i++) hx[(size_t)i] = x[i * incx];
  double* dx;
  GPC_CHECK(sc.alloc(&dx, (size_t)np, true));
  GPC_CHECK(up(sc, dx, np, hx.data(), n, n, 1));
  for (int64_t j0 = 0; j0 < n; j0 += 65535) {  // grid.y limit
    const int64_t nj = (n - j0 < 65535) ? n - j0 : 65535;
    dim3 grid((unsigned)((n + 1023) / 1024 < 64 ? (n + 1023) / 1024 : 64), (unsigned)nj);
    syr_lower_kernel<<<grid, 256, 0, sc.s>>>(dl + j0 + j0 * np, np, n - j0, alpha, dx + j0);
    GPC_CUDA_CHECK(cudaGetLastError());
    sc.launches++;
  }
  // back to the caller's triangle only (the other one is not referenced by dsyr_)
  double* dsrc = dl;
  if (is(uplo, 'U')) {
    double* dt;
    GPC_CHECK(sc.alloc(&dt, (size_t)np * np, true));
    GPC_CHECK(launch_transpose(dl, np, dt, np, n, n, sc.s, &sc.launches));
    dsrc = dt;
  }
  std::vector<double> h((size_t)n * n);
  GPC_CHECK(down(sc, h.data(), n, dsrc, np, n, n));
  for (int64_t j = 0; j < n; j++) {
    const int64_t lo = is(uplo, 'U') ? 0 : j, hi = is(uplo, 'U') ? j + 1 : n;
    for (int64_t i = lo; i < hi; i++) A[i + j * lda] = h[(size_t)(i + j * n)];
  }
  return GPC_OK;
}
extern "C" {

// ------------------------------------------------------------------------------------------------------
// measurement helpers
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i][0] = acc[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(acc[i][0]), "+d"(acc[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[0] = s;
}

int gpc_bench_dmma_peak(int device, double* tflops) {
  Scratch sc;
  GPC_CHECK(sc.init(device));
  cudaDeviceProp prop;
  GPC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  double* out;
  GPC_CHECK(sc.alloc(&out, 16, true));
  int ctas = prop.multiProcessorCount * 2, iters = 4096;
  cudaEvent_t e0, e1;
  GPC_CUDA_CHECK(cudaEventCreate(&e0));
  GPC_CUDA_CHECK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    GPC_CUDA_CHECK(cudaEventRecord(e0, sc.s));
    dmma_peak_kernel<<<ctas, 256, 0, sc.s>>>(out, iters);
    GPC_CUDA_CHECK(cudaEventRecord(e1, sc.s));
    GPC_CUDA_CHECK(cudaStreamSynchronize(sc.s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)ctas * 8 /*warps*/ * iters * 16.0 * 512.0;  // m8n8k4 = 256 FMA = 512 flop
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (tflops) *tflops = best;
  return GPC_OK;
}

int gpc_bench_gemm(int device, int64_t m, int64_t n, int64_t k, int a_kc, int b_kc, int lower, int cfg, int reps,
                   double* ms_out) {
  if (m % TILE || n % TILE || k % 16 || reps < 1) {
    set_error("gpc_bench_gemm: m, n must be multiples of 128 and k of 16");
    return GPC_ERR_ARG;
  }
  Scratch sc;
  GPC_CHECK(sc.init(device));
  double *A, *B, *C;
  GPC_CHECK(sc.alloc(&A, (size_t)m * k, true));
  GPC_CHECK(sc.alloc(&B, (size_t)n * k, true));
  GPC_CHECK(sc.alloc(&C, (size_t)m * n, true));
  GemmCall g{A, B, C, a_kc ? k : m, b_kc ? k : n, m, m, n, k, -1.0, 1.0, a_kc != 0, b_kc != 0, lower != 0};
  gemm_force_config(cfg);
  cudaEvent_t e0, e1;
  GPC_CUDA_CHECK(cudaEventCreate(&e0));
  GPC_CUDA_CHECK(cudaEventCreate(&e1));
  int rc = launch_gemm(g, sc.s, &sc.launches);  // warm-up
  if (rc == GPC_OK) {
    cudaEventRecord(e0, sc.s);
    for (int r = 0; r < reps && rc == GPC_OK; r++) rc = launch_gemm(g, sc.s, &sc.launches);
    cudaEventRecord(e1, sc.s);
    if (cudaStreamSynchronize(sc.s) != cudaSuccess) rc = GPC_ERR_CUDA;
  }
  gemm_force_config(-1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_out) *ms_out = ms / reps;
  return rc;
}

int gpc_set_gemm_engine(int ozaki, int slices, int64_t min_mn, int64_t min_k) {
  if (slices != 0 && (slices < 2 || slices > 8)) {
    set_error("gpc_set_gemm_engine: slices must be 2..8");
    return GPC_ERR_ARG;
  }
  oz_configure(ozaki, slices, min_mn, min_k);
  return GPC_OK;
}

int gpc_gemm_engine_slices(void) { return oz_slices(); }

// clock64 phase stamps (8 entries, relative to the CTA's start) of CTA `cta` of one tensor-core GEMM m x n x k:
// [0] start [1] barriers + TMEM ready [2] first operands landed [3] last MMA issued [4] C segment requested
// [5] accumulators complete [6] epilogue stores issued [7] end
int gpc_bench_oz_stamps(int device, int64_t m, int64_t n, int64_t k, int cta, long long* stamps8) {
  if (m % TILE || n % TILE || k % TILE || !stamps8) return GPC_ERR_ARG;
  Scratch sc;
  GPC_CHECK(sc.init(device));
  double *A, *B, *Cm;
  long long* st;
  GPC_CHECK(sc.alloc(&A, (size_t)m * k, true));
  GPC_CHECK(sc.alloc(&B, (size_t)n * k, true));
  GPC_CHECK(sc.alloc(&Cm, (size_t)m * n, true));
  GPC_CHECK(sc.alloc((double**)&st, 16, true));
  GemmCall g{A, B, Cm, m, n, m, m, n, k, -1.0, 1.0, false, false, false};
  GPC_CHECK(launch_gemm_ozaki(g, sc.s, &sc.launches));  // warm-up (workspaces, attributes)
  oz_set_debug(st, cta);
  int rc = launch_gemm_ozaki(g, sc.s, &sc.launches);
  oz_set_debug(nullptr, 0);
  if (rc != GPC_OK) return rc;
  long long h[8];
  GPC_CUDA_CHECK(cudaMemcpyAsync(h, st, sizeof(h), cudaMemcpyDeviceToHost, sc.s));
  GPC_CUDA_CHECK(cudaStreamSynchronize(sc.s));
  for (int i = 0; i < 8; i++) stamps8[i] = h[i] - h[0];
  return GPC_OK;
}

int gpc_bench_leaf(int device, int reps, double* us_out, long long* stamps_out) {
  // the 128 x 128 diagonal-block kernel on a well-conditioned SPD block, timed in isolation; stamps_out (32 entries)
  // receives the clock64 phase stamps of the last launch relative to its start
  if (reps < 1) return GPC_ERR_ARG;
  Scratch sc;
  GPC_CHECK(sc.init(device));
  double *A, *A0, *Dinv, *W, *ld;
  int* info;
  long long* st;
  GPC_CHECK(sc.alloc(&A, (size_t)TILE * TILE, true));
  GPC_CHECK(sc.alloc(&A0, (size_t)TILE * TILE, true));
  GPC_CHECK(sc.alloc(&Dinv, (size_t)TILE * TILE, true));
  GPC_CHECK(sc.alloc(&W, (size_t)TILE * TILE, true));
  GPC_CHECK(sc.alloc(&ld, 8, true));
  GPC_CHECK(sc.alloc((double**)&info, 8, true));
  GPC_CHECK(sc.alloc((double**)&st, 64, true));
  std::vector<double> h((size_t)TILE * TILE);
  for (int j = 0; j < TILE; j++)
    for (int i = 0; i < TILE; i++) h[i + (size_t)j * TILE] = (i == j ? 2.0 : 0.0) + exp(-0.05 * (i - j) * (i - j));
  GPC_CUDA_CHECK(cudaMemcpyAsync(A0, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, sc.s));
  cudaEvent_t e0, e1;
  GPC_CUDA_CHECK(cudaEventCreate(&e0));
  GPC_CUDA_CHECK(cudaEventCreate(&e1));
  float total = 0.f;
  for (int r = 0; r < reps + 1; r++) {
    GPC_CUDA_CHECK(cudaMemcpyAsync(A, A0, h.size() * sizeof(double), cudaMemcpyDeviceToDevice, sc.s));
    if (r == reps) leaf_set_debug(st);
    GPC_CUDA_CHECK(cudaEventRecord(e0, sc.s));
    int rc = launch_potrf_leaf(A, TILE, nullptr, info, 0, TILE, ld, sc.s, &sc.launches, W, TILE);
    leaf_set_debug(nullptr);
    if (rc != GPC_OK) return rc;
    GPC_CUDA_CHECK(cudaEventRecord(e1, sc.s));
    GPC_CUDA_CHECK(cudaStreamSynchronize(sc.s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && r < reps) total += ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (us_out) *us_out = 1e3 * total / (reps > 1 ? reps - 1 : 1);
  if (stamps_out) {
    long long hs[32];
    GPC_CUDA_CHECK(cudaMemcpy(hs, st, sizeof(hs), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 32; i++) stamps_out[i] = hs[i] ? hs[i] - hs[0] : 0;
  }
  return GPC_OK;
}

int gpc_gemm_check(int device, int64_t m, int64_t n, int64_t k, int a_kc, int b_kc, int lower, int cfg, double alpha,
                   double beta, const double* A, const double* B, double* C) {
  if (m % TILE || n % TILE || k % 128 || m < TILE || n < TILE || k < 128 || !A || !B || !C) {
    set_error("gpc_gemm_check: m, n, k must be positive multiples of 128");
    return GPC_ERR_ARG;
  }
  Scratch sc;
  GPC_CHECK(sc.init(device));
  double *dA, *dB, *dC;
  GPC_CHECK(sc.alloc(&dA, (size_t)m * k, false));
  GPC_CHECK(sc.alloc(&dB, (size_t)n * k, false));
  GPC_CHECK(sc.alloc(&dC, (size_t)m * n, false));
  GPC_CUDA_CHECK(cudaMemcpyAsync(dA, A, (size_t)m * k * sizeof(double), cudaMemcpyHostToDevice, sc.s));
  GPC_CUDA_CHECK(cudaMemcpyAsync(dB, B, (size_t)n * k * sizeof(double), cudaMemcpyHostToDevice, sc.s));
  GPC_CUDA_CHECK(cudaMemcpyAsync(dC, C, (size_t)m * n * sizeof(double), cudaMemcpyHostToDevice, sc.s));
  GemmCall g{dA, dB, dC, a_kc ? k : m, b_kc ? k : n, m, m, n, k, alpha, beta, a_kc != 0, b_kc != 0, (lower & 1) != 0};
  g.a_tri = ((lower >> 1) & 3) == 1 ? 1 : (((lower >> 1) & 3) == 2 ? -1 : 0);
  g.b_tri = ((lower >> 3) & 3) == 1 ? 1 : (((lower >> 3) & 3) == 2 ? -1 : 0);
  gemm_force_config(cfg);
  int rc = launch_gemm(g, sc.s, &sc.launches);
  gemm_force_config(-1);
  if (rc != GPC_OK) return rc;
  GPC_CUDA_CHECK(cudaMemcpyAsync(C, dC, (size_t)m * n * sizeof(double), cudaMemcpyDeviceToHost, sc.s));
  GPC_CUDA_CHECK(cudaStreamSynchronize(sc.s));
  return GPC_OK;
}

int gpc_bench_syrk(int device, int64_t n, int64_t k, int reps, double* ms_out) {
  if (n % TILE || k % 16 || n < TILE || k < 16 || reps < 1) {
    set_error("gpc_bench_syrk: n must be a multiple of 128 and k of 16");
    return GPC_ERR_ARG;
  }
  Scratch sc;
  GPC_CHECK(sc.init(device));
  double *A, *C;
  GPC_CHECK(sc.alloc(&A, (size_t)n * k, true));
  GPC_CHECK(sc.alloc(&C, (size_t)n * n, true));
  GemmCall g{A, A, C, n, n, n, n, n, k, -1.0, 1.0, false, false, true};
  cudaEvent_t e0, e1;
  GPC_CUDA_CHECK(cudaEventCreate(&e0));
  GPC_CUDA_CHECK(cudaEventCreate(&e1));
  GPC_CHECK(launch_gemm(g, sc.s, &sc.launches));  // warm-up
  GPC_CUDA_CHECK(cudaEventRecord(e0, sc.s));
  for (int r = 0; r < reps; r++) GPC_CHECK(launch_gemm(g, sc.s, &sc.launches));
  GPC_CUDA_CHECK(cudaEventRecord(e1, sc.s));
  GPC_CUDA_CHECK(cudaStreamSynchronize(sc.s));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_out) *ms_out = ms / reps;
  return GPC_OK;
}

}  // extern "C"
