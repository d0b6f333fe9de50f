// api.cu -- the C ABI of libgpc_b200.so (include/gpc_b200.h): device context, the recursive blocked
// Cholesky / triangular solves / SPD inverse built on the DMMA GEMM engine (dense.cu), the fused GP
// evaluation, and the lapack.h-level drop-ins.  Host code only orchestrates launches; there is no CPU
// fallback: every entry point needs a CUDA device.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <chrono>
#include "common.cuh"

// One evaluation uses ~20 streams (chain, bulk, side streams of the recursion).  With the default of 8 hardware work
// queues several streams share a queue and a small chain kernel can sit behind the queued launches of a bulk product
// on another stream (measured: 1 ms stalls, profiles/timeline_c2_r02.txt).  The driver reads the variable when the
// context is created, so it is set when the library is loaded -- without overriding a value the user chose.
__attribute__((constructor)) static void gpc_more_hw_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

namespace gpc {

static thread_local std::string g_error;
void set_error(const std::string& s) { g_error = s; }
int trace_sync(const char* what, cudaStream_t s) {
  static int on = -1;
  if (on < 0) on = getenv("GPC_TRACE") ? atoi(getenv("GPC_TRACE")) : 0;
  if (on < 2) return GPC_OK;
  fprintf(stderr, "[gpc trace] launched %s ...", what);
  fflush(stderr);
  cudaError_t e = cudaStreamSynchronize(s);
  fprintf(stderr, " %s\n", cudaGetErrorString(e));
  fflush(stderr);
  return e == cudaSuccess ? GPC_OK : GPC_ERR_CUDA;
}

#define GPC_CHECK(expr)             \
  do {                              \
    int _rc = (expr);               \
    if (_rc != GPC_OK) return _rc;  \
  } while (0)

// ------------------------------------------------------------------------------------------------------
// Recursive blocked algorithms.  All dimensions are multiples of TILE; recursion splits at multiples of TILE
// so that (almost) all flops are large-k GEMM/SYRK calls; the TILE x TILE leaves use the inverse of the
// factor's diagonal block (Dinv, written by the potrf leaf), turning the leaf TRSM into a GEMM as well.
// ------------------------------------------------------------------------------------------------------
static inline int64_t split(int64_t n) { return (n / TILE / 2) * TILE; }

// every GEMM of the recursion goes through here so that profiling mode can bracket it with events
// SMs a bulk product may hold at a time while the serial chain of the factorisation runs next to it (GemmCall::sm_limit)
static int bulk_sm_limit() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GPC_BULK_SMS");
    v = e ? atoi(e) : 132;
  }
  return v;
}

static int gemm(const Dense& d, const GemmCall& g0) {
  GemmCall g = g0;
  if (d.sm_limit > 0 && g.sm_limit == 0) g.sm_limit = d.sm_limit;
  if (!d.prof) return launch_gemm(g, d.s, d.launches);
  GemmProf* p = d.prof;
  if (p->used + 2 > p->ev.size()) {
    size_t old = p->ev.size();
    p->ev.resize(old + 1024);
    for (size_t i = old; i < p->ev.size(); i++) GPC_CUDA_CHECK(cudaEventCreate(&p->ev[i]));
  }
  GPC_CUDA_CHECK(cudaEventRecord(p->ev[p->used], d.s));
  int rc = launch_gemm(g, d.s, d.launches);
  GPC_CUDA_CHECK(cudaEventRecord(p->ev[p->used + 1], d.s));
  p->used += 2;
  double fl = g.lower ? (double)g.m * (double)(g.m + TILE) * (double)g.k : 2.0 * (double)g.m * (double)g.n * (double)g.k;
  // triangular operand: the zero k range is skipped (lower + triangular = the W'W product of the inverse, n^3/3)
  if (g.ktri || g.a_tri || g.b_tri) fl = g.lower ? (double)g.m * (double)g.m * (double)g.k / 3.0 : 0.5 * fl;
  p->flops.push_back(fl);
  const bool inplace = (g.C == g.A || g.C == g.B);
  p->recs.push_back({g.m, g.n, g.k, g.lower ? 1 : 0, (!inplace && oz_wants(g)) ? 1 : 0});
  return rc;
}

// X L' = B, in place.  B: m x n, L: n x n lower; dbase = index of L's first row/col in the full factor
int trsm_rlt(const Dense& d, double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl, int64_t n,
                    int64_t dbase) {
  if (m <= 0) return GPC_OK;
  if (n == TILE) {
    int64_t ldd;
    const double* Di = d.dinv_blk(dbase, &ldd);
    GemmCall g{B, Di, B, ldb, ldd, ldb, m, TILE, TILE, 1.0, 0.0, false, false, false};
    return gemm(d, g);  // force-128 rule inside launch_gemm keeps in-place safe (n == TILE)
  }
  int64_t n1 = split(n), n2 = n - n1;
  GPC_CHECK(trsm_rlt(d, B, ldb, m, L, ldl, n1, dbase));
  GemmCall g{B, L + n1, B + n1 * ldb, ldb, ldl, ldb, m, n2, n1, -1.0, 1.0, false, false, false};
  GPC_CHECK(gemm(d, g));
  return trsm_rlt(d, B + n1 * ldb, ldb, m, L + n1 + n1 * ldl, ldl, n2, dbase + n1);
}

// X L = B, in place.
int trsm_rln(const Dense& d, double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl, int64_t n,
                    int64_t dbase) {
  if (m <= 0) return GPC_OK;
  if (n == TILE) {
    int64_t ldd;
    const double* Di = d.dinv_blk(dbase, &ldd);
    GemmCall g{B, Di, B, ldb, ldd, ldb, m, TILE, TILE, 1.0, 0.0, false, true, false};
    return gemm(d, g);
  }
  int64_t n1 = split(n), n2 = n - n1;
  GPC_CHECK(trsm_rln(d, B + n1 * ldb, ldb, m, L + n1 + n1 * ldl, ldl, n2, dbase + n1));
  GemmCall g{B + n1 * ldb, L + n1, B, ldb, ldl, ldb, m, n1, n2, -1.0, 1.0, false, true, false};
  GPC_CHECK(gemm(d, g));
  return trsm_rln(d, B, ldb, m, L, ldl, n1, dbase);
}

// in-place lower Cholesky of A (n x n), only the lower triangle is referenced / written.
// Look-ahead: the trailing update A22 -= A21 A21' is issued in two parts -- the top-left block that the next diagonal
// factorisation needs on the main stream, the rest on a side stream -- and the recursion on A22 only waits for the
// side stream (`pending`) when it first touches rows below its own first block.
int potrf_rec(const Dense& d, double* A, int64_t lda, int64_t n, int64_t base, cudaEvent_t pending) {
  if (n == TILE) {
    if (pending) GPC_CUDA_CHECK(cudaStreamWaitEvent(d.s, pending, 0));
    return launch_potrf_leaf(A, lda, d.Dinv + base * TILE, d.info, (int)base, d.nvalid - base, d.logdet, d.s,
                             d.launches);
  }
  int64_t n1 = split(n), n2 = n - n1;
  GPC_CHECK(potrf_rec(d, A, lda, n1, base, nullptr));
  if (pending) GPC_CUDA_CHECK(cudaStreamWaitEvent(d.s, pending, 0));
  GPC_CHECK(trsm_rlt(d, A + n1, lda, n2, A, lda, n1, base));
  double* A21 = A + n1;
  double* A22 = A + n1 + n1 * lda;
  if (d.fk && n2 > TILE) {
    int64_t m1 = split(n2), m2 = n2 - m1;
    cudaEvent_t e_trsm = d.fk->event(), e_side = d.fk->event();
    cudaStream_t side = d.fk->stream();
    GPC_CUDA_CHECK(cudaEventRecord(e_trsm, d.s));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(side, e_trsm, 0));
    Dense ds = d;
    ds.s = side;
    {  // main: A22[0:m1,0:m1] -= A21[0:m1,:] A21[0:m1,:]'
      GemmCall g{A21, A21, A22, lda, lda, lda, m1, m1, n1, -1.0, 1.0, false, false, true};
      GPC_CHECK(gemm(d, g));
    }
    {  // side: A22[m1:,0:m1] -= A21[m1:,:] A21[0:m1,:]'   and   A22[m1:,m1:] -= A21[m1:,:] A21[m1:,:]'
      GemmCall g1{A21 + m1, A21, A22 + m1, lda, lda, lda, m2, m1, n1, -1.0, 1.0, false, false, false};
      GPC_CHECK(gemm(ds, g1));
      GemmCall g2{A21 + m1, A21 + m1, A22 + m1 + m1 * lda, lda, lda, lda, m2, m2, n1, -1.0, 1.0, false, false, true};
      GPC_CHECK(gemm(ds, g2));
    }
    GPC_CUDA_CHECK(cudaEventRecord(e_side, side));
    return potrf_rec(d, A22, lda, n2, base + n1, e_side);
  }
  GemmCall g{A21, A21, A22, lda, lda, lda, n2, n2, n1, -1.0, 1.0, false, false, true};
  GPC_CHECK(gemm(d, g));
  return potrf_rec(d, A22, lda, n2, base + n1, nullptr);
}

size_t potri_workspace(int64_t n) {
  if (n <= TILE) return 0;
  int64_t n1 = split(n), n2 = n - n1;
  return (size_t)n1 * n2 + potri_workspace(n1) + potri_workspace(n2);
}

// Out (n x n, full symmetric) = (L L')^-1 by the Schur-complement recursion:
//   K^-1 = [A^-1 + X' S^-1 X, -X' S^-1; -S^-1 X, S^-1],  X = L21 L11^-1, S^-1 = (L22 L22')^-1
// The three sub-problems (A^-1, S^-1, X) are independent: with a Fork they run on separate streams.
int potri_rec(const Dense& d, const double* L, int64_t ldl, int64_t n, double* Out, int64_t ldo, int64_t dbase,
              double* W) {
  if (!W) W = d.W;
  if (n == TILE) {
    const double* Di = d.Dinv + dbase * TILE;
    GemmCall g{Di, Di, Out, TILE, TILE, ldo, TILE, TILE, TILE, 1.0, 0.0, true, true, false};
    return gemm(d, g);
  }
  int64_t n1 = split(n), n2 = n - n1;
  double* X = W;  // n2 x n1, ld n2
  double* W1 = W + n1 * n2;
  double* W2 = W1 + potri_workspace(n1);
  if (d.fk && n >= 4 * TILE) {
    cudaEvent_t e0 = d.fk->event(), e1 = d.fk->event(), e2 = d.fk->event();
    Dense d1 = d, d2 = d;
    d1.s = d.fk->stream();
    d2.s = d.fk->stream();
    GPC_CUDA_CHECK(cudaEventRecord(e0, d.s));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(d1.s, e0, 0));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(d2.s, e0, 0));
    GPC_CHECK(potri_rec(d1, L, ldl, n1, Out, ldo, dbase, W1));
    GPC_CHECK(potri_rec(d2, L + n1 + n1 * ldl, ldl, n2, Out + n1 + n1 * ldo, ldo, dbase + n1, W2));
    GPC_CHECK(launch_copy_block(L + n1, ldl, X, n2, n2, n1, 1.0, d.s, d.launches));
    GPC_CHECK(trsm_rln(d, X, n2, n2, L, ldl, n1, dbase));
    GPC_CUDA_CHECK(cudaEventRecord(e1, d1.s));
    GPC_CUDA_CHECK(cudaEventRecord(e2, d2.s));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(d.s, e1, 0));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(d.s, e2, 0));
  } else {
    GPC_CHECK(potri_rec(d, L, ldl, n1, Out, ldo, dbase, W1));
    GPC_CHECK(potri_rec(d, L + n1 + n1 * ldl, ldl, n2, Out + n1 + n1 * ldo, ldo, dbase + n1, W2));
    GPC_CHECK(launch_copy_block(L + n1, ldl, X, n2, n2, n1, 1.0, d.s, d.launches));
    GPC_CHECK(trsm_rln(d, X, n2, n2, L, ldl, n1, dbase));
  }
  {  // Out21 = -Out22 * X
    GemmCall g{Out + n1 + n1 * ldo, X, Out + n1, ldo, n2, ldo, n2, n1, n2, -1.0, 0.0, false, true, false};
    GPC_CHECK(gemm(d, g));
  }
  {  // Out11 -= X' * Out21 (lower), then mirror
    GemmCall g{X, Out + n1, Out, n2, ldo, ldo, n1, n1, n2, -1.0, 1.0, true, true, true};
    GPC_CHECK(gemm(d, g));
    GPC_CHECK(launch_mirror_lower(Out, ldo, n1, d.s, d.launches));
  }
  return launch_transpose(Out + n1, ldo, Out + n1 * ldo, ldo, n2, n1, d.s, d.launches);
}


// ------------------------------------------------------------------------------------------------------
// Factor + explicit triangular inverse.  W = L^-1 is built alongside L, so every panel solve is ONE large GEMM
// (L21 = A21 W11', triangular k-range skipped) instead of a recursion down to TILE-wide in-place leaves, and the
// inverse (K^-1 = W'W) is ONE triangular product.  Same flop count as dpotrf_ + dpotri_ (N^3/3 + 2N^3/3,
// reference lapack.h:59-73), but all of it in calls large enough for the tensor-core engine.
//   node(A, n):  node(A11) -> L11, W11
//                L21 = A21 W11'                 (main stream, via tmpL: the product cannot be formed in place)
//                T   = L21 W11                  (side stream, needed only for W21; a large one on the bulk stream)
//                A22 -= L21 L21'                (SYRK)
//                node(A22) -> L22, W22
//                W21 = -W22 T
// Inside gpc_eval (N <= 16384) the stream this runs on is the context's high-priority chain stream and d.bulk its
// low-priority bulk stream (wave-limited products, GemmCall::sm_limit); two optional restructurings of the TOP level hang
// off hooks in the recursion: TopFront (look-ahead of the a-columns of L21, on by default) and TopPipe (row-block
// pipeline of W21 / K^-1, off by default) -- DESIGN.md 4.3 "How one evaluation is scheduled".
// ------------------------------------------------------------------------------------------------------
size_t potrf_inv_tspace(int64_t n) {
  if (n <= TILE) return 0;
  int64_t n1 = split(n), n2 = n - n1;
  size_t a = potrf_inv_tspace(n1), b = (size_t)n1 * n2 + potrf_inv_tspace(n2);
  return a > b ? a : b;
}

static int winv_offdiag(const Dense& d, const double* A, int64_t lda, int64_t n1, int64_t n2, int64_t base, double* T,
                        bool t_done, cudaEvent_t t_ready) {
  double* W = d.Winv + base + base * d.ldw;
  if (!t_done) {  // T = L21 W11 (W11 lower: k starts at the tile's first column)
    GemmCall gt{A + n1, W, T, lda, d.ldw, n2, n2, n1, n1, 1.0, 0.0, false, true, false};
    gt.b_tri = +1;
    GPC_CHECK(gemm(d, gt));
  } else if (t_ready) {
    GPC_CUDA_CHECK(cudaStreamWaitEvent(d.s, t_ready, 0));
  }
  // W21 = -W22 T (W22 lower: k ends at the tile's last row)
  GemmCall gw{W + n1 + n1 * d.ldw, T, W + n1, d.ldw, n2, d.ldw, n2, n1, n2, -1.0, 0.0, false, true, false};
  gw.a_tri = -1;
  return gemm(d, gw);
}

int potrf_inv_rec(const Dense& d, double* A, int64_t lda, int64_t n, int64_t base, double* T, bool defer_top,
                  size_t tl_off) {
  // at the top level (defer_top) W21 waits for the inverse; its first factor T = L21 W11 is still started here on a
  // side stream when the caller announced that the inverse follows (d.top_t_ready): one full-GPU product that fills the
  // idle SMs of the latency-bound A22 recursion
  const bool eager_t = !defer_top || (d.top_t_ready != nullptr);
  double* W = d.Winv + base + base * d.ldw;
  if (n == TILE)
    return launch_potrf_leaf(A, lda, nullptr, d.info, (int)base, d.nvalid - base, d.logdet, d.s, d.launches, W, d.ldw);
  int64_t n1 = split(n), n2 = n - n1;
  double* A21 = A + n1;
  double* A22 = A + n1 + n1 * lda;
  // L21 = A21 W11' cannot be formed in place.  With side streams the copy of A21 (final on entry: every earlier
  // update is ordered before this point on the main stream) is taken NOW, concurrently with the A11 recursion, into
  // this depth's slot of the copy pool; the product then reads the copy and writes A21 directly.
  double* TL = d.TLpool ? d.TLpool + tl_off : nullptr;
  cudaEvent_t e_copy = nullptr;
  if (d.fk && TL) {
    cudaEvent_t e0 = d.fk->event();
    e_copy = d.fk->event();
    cudaStream_t side = d.fk->copy_stream();
    GPC_CUDA_CHECK(cudaEventRecord(e0, d.s));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(side, e0, 0));
    GPC_CHECK(launch_copy_block(A21, lda, TL, n2, n2, n1, 1.0, side, d.launches));
    GPC_CUDA_CHECK(cudaEventRecord(e_copy, side));
  }
  TopFront* tf = (d.tf && d.tf->on && d.bulk) ? d.tf : nullptr;
  if (tf && defer_top) {  // announce the top-level panel to the A11 node (hook below)
    tf->queued = false;
    if (e_copy && n1 >= 8 * TILE) {
      tf->n1 = n1;
      tf->n2 = n2;
      tf->lda = lda;
      tf->A21 = A21;
      tf->A22 = A22;
      tf->TL = TL;
      tf->e_copy = e_copy;
    } else {
      tf->n1 = 0;
    }
  }
  GPC_CHECK(potrf_inv_rec(d, A, lda, n1, base, T, false, tl_off + (size_t)n1 * n2));
  if (tf && !defer_top && base == 0 && tf->n1 > 0 && n == tf->n1) {
    // the A11 node of the top level, its first diagonal node a (h = n1 rows) done:
    //   X1  L21[:, a] = A21[:, a] W_aa'        X2  A22 -= L21[:, a] L21[:, a]'       on the bulk stream, wave-limited
    tf->h = n1;
    GPC_CUDA_CHECK(cudaEventRecord(tf->e_a, d.s));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(d.bulk, tf->e_a, 0));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(d.bulk, tf->e_copy, 0));
    Dense db = d;
    db.s = d.bulk;
    db.sm_limit = bulk_sm_limit();
    {
      GemmCall g{tf->TL, d.Winv, tf->A21, tf->n2, d.ldw, tf->lda, tf->n2, n1, n1, 1.0, 0.0, false, false, false};
      g.b_tri = -1;
      GPC_CHECK(gemm(db, g));
    }
    {
      GemmCall g{tf->A21, tf->A21, tf->A22, tf->lda, tf->lda, tf->lda, tf->n2, tf->n2, n1, -1.0, 1.0, false, false, true};
      GPC_CHECK(gemm(db, g));
    }
    GPC_CUDA_CHECK(cudaEventRecord(tf->e_x2, d.bulk));
    tf->queued = true;
  }
  const bool front = tf && defer_top && tf->queued;
  TopPipe* tp = (d.tp && d.tp->on) ? d.tp : nullptr;
  if (tp && defer_top) {
    // row block 1 (W11) is complete: its contribution K^-1_11 = W11' W11 runs on the bulk stream under the rest
    tp->n1 = n1;
    tp->n2 = n2;
    tp->y_queued = false;
    GPC_CUDA_CHECK(cudaEventRecord(tp->e_w11, d.s));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(tp->bulk, tp->e_w11, 0));
    Dense db = d;
    db.s = tp->bulk;
    db.sm_limit = bulk_sm_limit();
    GemmCall g{W, W, tp->Kinv, d.ldw, d.ldw, tp->ldo, n1, n1, n1, 1.0, 0.0, true, true, true};
    g.a_tri = +1;
    GPC_CHECK(gemm(db, g));
  } else if (tp && !defer_top && base == tp->n1 && n == tp->n2 && tp->t_ready) {
    // second half of the top level, first diagonal node a' done (L and W of rows base .. base + n1):
    //   Y1  W21[a', :] = -W_a'a' T[a', :]                                   (the top-level W21 is deferred: its rows a' now)
    //   Y3  K^-1[<= a', <= a'] += W[a', <= a']' W[a', <= a']                  (three products: 11 block, a' row, a'a' block)
    tp->h1 = n1;
    const int64_t t1 = tp->n1, t2 = tp->n2, h1 = n1;
    GPC_CUDA_CHECK(cudaEventRecord(tp->e_half, d.s));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(tp->bulk, tp->e_half, 0));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(tp->bulk, tp->t_ready, 0));
    Dense db = d;
    db.s = tp->bulk;
    db.sm_limit = bulk_sm_limit();
    double* Waa = d.Winv + t1 + t1 * d.ldw;  // W_a'a'
    double* W21a = d.Winv + t1;              // rows a' of the top-level W21
    {
      GemmCall g{Waa, tp->T, W21a, d.ldw, t2, d.ldw, h1, t1, h1, -1.0, 0.0, false, true, false};
      g.a_tri = -1;
      GPC_CHECK(gemm(db, g));
    }
    {
      GemmCall g{W21a, W21a, tp->Kinv, d.ldw, d.ldw, tp->ldo, t1, t1, h1, 1.0, 1.0, true, true, true};
      GPC_CHECK(gemm(db, g));
    }
    {
      GemmCall g{Waa, W21a, tp->Kinv + t1, d.ldw, d.ldw, tp->ldo, h1, t1, h1, 1.0, 0.0, true, true, false};
      g.a_tri = +1;
      GPC_CHECK(gemm(db, g));
    }
    {
      GemmCall g{Waa, Waa, tp->Kinv + t1 + t1 * tp->ldo, d.ldw, d.ldw, tp->ldo, h1, h1, h1, 1.0, 0.0, true, true, true};
      g.a_tri = +1;
      GPC_CHECK(gemm(db, g));
    }
    GPC_CUDA_CHECK(cudaEventRecord(tp->e_bulk, tp->bulk));
    tp->y_queued = true;
  }
  if (front) {
    // the a-columns of L21 are (being) done on the bulk stream; the b-columns now:
    //   L21[:, b] = A21[:, a] W_ba' + A21[:, b] W_bb'
    const int64_t h = tf->h, hb = n1 - h;
    GPC_CUDA_CHECK(cudaStreamWaitEvent(d.s, e_copy, 0));
    {
      GemmCall g{TL, d.Winv + h, A21 + h * lda, n2, d.ldw, lda, n2, hb, h, 1.0, 0.0, false, false, false};
      GPC_CHECK(gemm(d, g));
    }
    {
      GemmCall g{TL + h * n2, d.Winv + h + h * d.ldw, A21 + h * lda, n2, d.ldw, lda, n2, hb, hb, 1.0, 1.0, false, false, false};
      g.b_tri = -1;
      GPC_CHECK(gemm(d, g));
    }
  } else if (e_copy) {  // W11'(kk, j) = W11(j, kk) is zero for kk > j: k ends at the tile's last column
    GPC_CUDA_CHECK(cudaStreamWaitEvent(d.s, e_copy, 0));
    GemmCall g{TL, W, A21, n2, d.ldw, lda, n2, n1, n1, 1.0, 0.0, false, false, false};
    g.b_tri = -1;
    GPC_CHECK(gemm(d, g));
  } else {
    GemmCall g{A21, W, d.tmpL, lda, d.ldw, n2, n2, n1, n1, 1.0, 0.0, false, false, false};
    g.b_tri = -1;
    GPC_CHECK(gemm(d, g));
    GPC_CHECK(launch_copy_block(d.tmpL, n2, A21, lda, n2, n1, 1.0, d.s, d.launches));
  }
  cudaEvent_t t_ready = nullptr;
  bool t_done = false;
  if (eager_t && d.fk) {  // T on a side stream, concurrent with the SYRK and the A22 recursion
    cudaEvent_t e1 = d.fk->event();
    t_ready = d.fk->event();
    Dense ds = d;
    GemmCall gt{A21, W, T, lda, d.ldw, n2, n2, n1, n1, 1.0, 0.0, false, true, false};
    gt.b_tri = +1;
    // T is needed only for W21: it runs next to the chain of the A22 recursion.  A large one (tensor-core engine) goes
    // to the ONE low-priority bulk stream, wave-limited so that the chain always finds free SMs; a small one to a
    // (high-priority) side stream
    if (d.bulk && oz_wants(gt)) {
      ds.s = d.bulk;
      ds.sm_limit = bulk_sm_limit();
    } else {
      ds.s = d.fk->stream();
    }
    GPC_CUDA_CHECK(cudaEventRecord(e1, d.s));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(ds.s, e1, 0));
    GPC_CHECK(gemm(ds, gt));
    GPC_CUDA_CHECK(cudaEventRecord(t_ready, ds.s));
    t_done = true;
    if (tp && defer_top) {
      tp->t_ready = t_ready;
      tp->T = T;
    }
  }
  if (front) {  // the a-columns' share of the update is on the bulk stream: wait for it, then the b-columns' share
    const int64_t h = tf->h, hb = n1 - h;
    GPC_CUDA_CHECK(cudaStreamWaitEvent(d.s, tf->e_x2, 0));
    GemmCall g{A21 + h * lda, A21 + h * lda, A22, lda, lda, lda, n2, n2, hb, -1.0, 1.0, false, false, true};
    GPC_CHECK(gemm(d, g));
    tf->queued = false;
  } else {
    GemmCall g{A21, A21, A22, lda, lda, lda, n2, n2, n1, -1.0, 1.0, false, false, true};
    GPC_CHECK(gemm(d, g));
  }
  GPC_CHECK(potrf_inv_rec(d, A22, lda, n2, base + n1, T + (size_t)n1 * n2, false, tl_off + (size_t)n1 * n2));
  if (defer_top) {
    if (d.top_t_ready) *d.top_t_ready = t_done ? t_ready : nullptr;
    return GPC_OK;
  }
  return winv_offdiag(d, A, lda, n1, n2, base, T, t_done, t_ready);
}

// the deferred top-level W21 (no-op when nothing was deferred): after this W = L^-1 is complete
int complete_W(const Dense& d, const double* L, int64_t ldl, int64_t n, bool deferred_top) {
  if (deferred_top && d.tp && d.tp->on && d.tp->y_queued) {
    // rows a' of the top-level W21 were done under the factorisation of b' (TopPipe); rows b' now:
    //   W21[b', :] = -(W_b'a' T[a', :] + W_b'b' T[b', :])
    const TopPipe& tp = *d.tp;
    const int64_t t1 = tp.n1, t2 = tp.n2, h1 = tp.h1, h2 = t2 - h1;
    GPC_CUDA_CHECK(cudaStreamWaitEvent(d.s, tp.t_ready, 0));
    double* C = d.Winv + t1 + h1;
    {
      GemmCall g{d.Winv + (t1 + h1) + t1 * d.ldw, tp.T, C, d.ldw, t2, d.ldw, h2, t1, h1, -1.0, 0.0, false, true, false};
      GPC_CHECK(gemm(d, g));
    }
    {
      GemmCall g{d.Winv + (t1 + h1) + (t1 + h1) * d.ldw, tp.T + h1, C, d.ldw, t2, d.ldw, h2, t1, h2, -1.0, 1.0, false, true, false};
      g.a_tri = -1;
      GPC_CHECK(gemm(d, g));
    }
    if (d.top_t_ready) *d.top_t_ready = nullptr;
    return GPC_OK;
  }
  if (deferred_top && n > TILE) {
    int64_t n1 = split(n), n2 = n - n1;
    const cudaEvent_t t_ready = d.top_t_ready ? *d.top_t_ready : nullptr;  // T already queued by the factorisation?
    GPC_CHECK(winv_offdiag(d, L, ldl, n1, n2, 0, d.Tpool, t_ready != nullptr, t_ready));
    if (d.top_t_ready) *d.top_t_ready = nullptr;
  }
  return GPC_OK;
}

// Out (lower triangle; mirrored into the upper one when `mirror`) = W' W = (L L')^-1 from the complete W
int kinv_from_W(const Dense& d, int64_t n, double* Out, int64_t ldo, bool mirror) {
  if (d.tp && d.tp->on && d.tp->y_queued && Out == d.tp->Kinv) {
    // the contributions of the row blocks 1 and a' are (being) accumulated on the bulk stream; row block b' now:
    //   K^-1[< b', < b'] += R'R,  K^-1[b', < b'] = W_b'b'' R,  K^-1[b', b'] = W_b'b'' W_b'b',   R = W[b', < b']
    TopPipe& tp = *d.tp;
    const int64_t t1 = tp.n1, h1 = tp.h1, h2 = tp.n2 - h1, m = t1 + h1;
    GPC_CUDA_CHECK(cudaStreamWaitEvent(d.s, tp.e_bulk, 0));
    double* R = d.Winv + m;                  // rows b', columns 0 .. m-1
    double* Wbb = d.Winv + m + m * d.ldw;    // W_b'b'
    {
      GemmCall g{R, R, Out, d.ldw, d.ldw, ldo, m, m, h2, 1.0, 1.0, true, true, true};
      GPC_CHECK(gemm(d, g));
    }
    {
      GemmCall g{Wbb, R, Out + m, d.ldw, d.ldw, ldo, h2, m, h2, 1.0, 0.0, true, true, false};
      g.a_tri = +1;
      GPC_CHECK(gemm(d, g));
    }
    {
      GemmCall g{Wbb, Wbb, Out + m + m * ldo, d.ldw, d.ldw, ldo, h2, h2, h2, 1.0, 0.0, true, true, true};
      g.a_tri = +1;
      GPC_CHECK(gemm(d, g));
    }
    tp.y_queued = false;
    return mirror ? launch_mirror_lower(Out, ldo, n, d.s, d.launches) : GPC_OK;
  }
  // Out(i, j) = sum_{kk >= i} W(kk, i) W(kk, j), i >= j: lower tiles only, k starts at the tile's first row
  GemmCall g{d.Winv, d.Winv, Out, d.ldw, d.ldw, ldo, n, n, n, 1.0, 0.0, true, true, true};
  g.a_tri = +1;
  GPC_CHECK(gemm(d, g));
  return mirror ? launch_mirror_lower(Out, ldo, n, d.s, d.launches) : GPC_OK;
}

int inverse_from_W(const Dense& d, const double* L, int64_t ldl, int64_t n, double* Out, int64_t ldo, bool deferred_top) {
  GPC_CHECK(complete_W(d, L, ldl, n, deferred_top));
  return kinv_from_W(d, n, Out, ldo, true);
}

}  // namespace gpc

using namespace gpc;

// ------------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------------
struct gpc_ctx {
  int device;
  cudaStream_t own_stream, stream;
  int64_t Nmax, Npmax;
  int Dmax, dmax;
  int64_t N, Np;
  int D, d;
  double *X, *M, *alpha, *K, *L, *Kinv, *W, *Dinv;
  double* Winv;    // W = L^-1 (lower block triangle), built by potrf_inv_rec; Kinv doubles as its scratch (tmpL, Tpool)
  bool w_deferred; // the top-level W21 has not been formed yet
  bool inverse_follows;        // set by gpc_eval: the factorisation may start the top-level T early
  cudaEvent_t top_t_ready;     // non-null: T of the top level is being computed on a side stream
  int64_t winv_np; // padded size for which the diagonal blocks of Winv have zero strict upper parts (0: never)
  bool use_winv;   // GPC_POTRF_MODE != "rec": factor + explicit inverse (default); "rec": recursive TRSM / Schur inverse
  double* scal;    // device scalars: [0] logdet [1] quad [2] trace ; then g[GPC_MAX_PARAMS]
  int* info;       // device
  double* partial; // grad partial sums
  double* gXdev;
  double* symm_part;  // scratch of the alpha = K^-1 m product
  int max_ctas;
  double* hres;    // pinned host result buffer
  int* hinfo;      // pinned
  bool haveX, haveM, haveK, haveL, haveInv, haveAlpha;
  // gpc_eval builds K straight into the buffer that is factored in place; the K buffer itself is rebuilt on demand
  // (gpc_download, gpc_potrf, ...) from the kernel of that evaluation
  bool k_lazy;
  KSpec ks_last;
  bool kinv_full;      // the upper triangle of K^-1 is a mirror of the lower one (gpc_eval leaves it unwritten)
  double *zw, *trmv_part;  // alpha = W'(W m): the intermediate vector and the partial sums of the two products
  // gpc_eval: the factorisation runs on a high-priority stream (its serial chain of small kernels must not queue behind
  // the CTAs of the side-stream products), the row-block pipeline of the top level on a low-priority one
  cudaStream_t chain_stream, bulk_stream;
  cudaEvent_t ev_pipe[6];   // chain fork / join, W11 ready, a' ready, bulk done, bulk joined
  TopPipe pipe;
  TopFront front;
  cudaEvent_t ev_front[2];
  double* pipe_scratch;     // tmpL / Tpool / TLpool outside K^-1 (the pipeline writes K^-1 while they are in use)
  int64_t launches;
  cudaEvent_t ev[6];
  double last_ms[6];
  // scratch for cross-covariances (grown on demand)
  double *Xs, *Kc, *Kc2, *tmp1, *tmp2;
  int64_t Xs_cap, Kc_cap, tmp_cap;
  Fork* fork;           // side streams / events for the recursions
  GemmProf* prof;       // non-null in profiling mode
  double prof_ms, prof_flops;
  int64_t prof_count;
  double enqueue_ms;     // host time gpc_eval spent queueing the last evaluation (before its single synchronisation)
  double prof_split[8];  // DMMA [ms, flops, launches, 0], Ozaki [ms, fp64-equivalent flops, launches, int8 ops]
};

static const int SC_LOGDET = 0, SC_QUAD = 1, SC_TRACE = 2, SC_G = 8;

static Dense dense_of(gpc_ctx* c) {
  Dense d;
  d.prof = c->prof;
  d.fk = c->prof ? nullptr : c->fork;  // profiling mode runs serially so that per-launch durations are not overlapped
  d.s = c->stream;
  d.launches = &c->launches;
  d.Dinv = c->Dinv;
  d.info = c->info;
  d.logdet = c->scal + SC_LOGDET;
  d.W = c->W;
  d.nvalid = c->N;
  d.Winv = c->Winv;
  d.top_t_ready = nullptr;
  d.ldw = c->Np;
  d.tp = &c->pipe;
  d.tf = &c->front;
  d.bulk = (c->stream == c->chain_stream) ? c->bulk_stream : nullptr;  // only while gpc_eval runs the chain at priority
  // scratch inside the (not yet written) K^-1 buffer: tmpL then Tpool -- or in its own buffer when the row-block
  // pipeline may write K^-1 during the factorisation
  double* scr = c->pipe_scratch ? c->pipe_scratch : c->Kinv;
  d.tmpL = scr;
  d.Tpool = scr ? scr + (size_t)(c->Npmax / 2 + TILE) * (size_t)(c->Npmax / 2 + TILE) : nullptr;
  d.TLpool = d.Tpool ? d.Tpool + potrf_inv_tspace(c->Npmax) + 16 : nullptr;
  return d;
}

static int ensure_inverse_buffers(gpc_ctx* c) {
  // each buffer is checked on its own: a call that ran out of memory half-way is completed (or fails again) next time
  size_t nn = (size_t)c->Npmax * c->Npmax;
  // K^-1 also serves as scratch of potrf_inv_rec: tmpL ((Np/2+TILE)^2) + Tpool (<= Np^2/3 + slack)
  size_t scratch = (size_t)(c->Npmax / 2 + TILE) * (size_t)(c->Npmax / 2 + TILE) + 2 * (potrf_inv_tspace(c->Npmax) + 16);
  if (!c->Kinv) GPC_CUDA_CHECK(cudaMalloc(&c->Kinv, (nn > scratch ? nn : scratch) * sizeof(double)));
  if (c->use_winv && !c->Winv) {
    GPC_CUDA_CHECK(cudaMalloc(&c->Winv, nn * sizeof(double)));
    c->winv_np = 0;
  }
  if (!c->use_winv && !c->W) GPC_CUDA_CHECK(cudaMalloc(&c->W, (potri_workspace(c->Npmax) + 16) * sizeof(double)));
  if (!c->symm_part)
    GPC_CUDA_CHECK(cudaMalloc(&c->symm_part, (size_t)symm_chunks(c->Npmax) * 4 * c->Npmax * sizeof(double)));
  {
    // off by default: measured on B200 at C2 (profiles/timeline_c2_r02.txt) the bulk products slow the chain's own
    // mid-size products by more than the shorter tail returns (13.4 vs 12.4 ms); GPC_TOP_PIPE=1 switches it on
    const char* e = getenv("GPC_TOP_PIPE");
    const bool want = c->use_winv && c->fork && c->Npmax >= 1024 && c->Npmax <= 16384 && e && atoi(e) != 0;
    if (want && !c->pipe_scratch) GPC_CUDA_CHECK(cudaMalloc(&c->pipe_scratch, scratch * sizeof(double)));
  }
  if (!c->zw) GPC_CUDA_CHECK(cudaMalloc(&c->zw, (size_t)c->Npmax * c->dmax * sizeof(double)));
  if (!c->trmv_part) GPC_CUDA_CHECK(cudaMalloc(&c->trmv_part, (size_t)8 * 4 * c->Npmax * sizeof(double)));
  return GPC_OK;
}

extern "C" {

const char* gpc_last_error(void) { return g_error.c_str(); }

int gpc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int gpc_kern_nparams(int type, int D) {
  switch (type) {
    case GPC_KERN_WHITE:
    case GPC_KERN_BIAS:
    case GPC_KERN_LIN: return 1;
    case GPC_KERN_RBF:
    case GPC_KERN_MATERN32:
    case GPC_KERN_MATERN52: return 2;
    case GPC_KERN_POLY: return 3;
    case GPC_KERN_RBFARD: return 2 + D;
  }
  return -1;
}
int gpc_kern_transform(int type, int idx) {
  if (type == GPC_KERN_RBFARD && idx >= 2) return GPC_TRANS_SIGMOID;  // CKern.cpp:3210-3217 defaultZeroOne
  return GPC_TRANS_EXP;                                                // defaultPositive
}
static const double LIMVAL = 36.0;  // CTransform.h:17
double gpc_transform_atox(int tr, double a) {
  switch (tr) {
    case GPC_TRANS_EXP:  // CTransform.cpp:31-42
      if (a < -LIMVAL) return exp(-LIMVAL);
      if (a < LIMVAL) return exp(a);
      return exp(LIMVAL);
    case GPC_TRANS_SIGMOID:  // CTransform.cpp:96-104
      if (a < -LIMVAL) return 2.220446049250313e-16;
      if (a < LIMVAL) return 1.0 / (1.0 + exp(-a));
      return 1.0 - 2.220446049250313e-16;
  }
  return a;
}
double gpc_transform_xtoa(int tr, double x) {
  switch (tr) {
    case GPC_TRANS_EXP: return log(x);
    case GPC_TRANS_SIGMOID: return log(x / (1.0 - x));
  }
  return x;
}
double gpc_transform_gradfact(int tr, double x) {
  switch (tr) {
    case GPC_TRANS_EXP: return x;
    case GPC_TRANS_SIGMOID: return x * (1.0 - x);
  }
  return 1.0;
}

static int ctx_create_impl(gpc_ctx* c, int device, int64_t Nmax, int Dmax, int dout_max, const cudaDeviceProp& prop);

int gpc_ctx_create(gpc_ctx** out, int device, int64_t Nmax, int Dmax, int dout_max) {
  if (!out || Nmax < 1 || Dmax < 1 || dout_max < 1) {
    set_error("gpc_ctx_create: bad arguments");
    return GPC_ERR_ARG;
  }
  int ndev = 0;
  GPC_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) {
    set_error("gpc_ctx_create: no such CUDA device");
    return GPC_ERR_CUDA;
  }
  GPC_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  GPC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("gpc_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
    return GPC_ERR_CUDA;
  }
  gpc_ctx* c = new gpc_ctx();
  memset(c, 0, sizeof(*c));
  *out = nullptr;
  int rc = ctx_create_impl(c, device, Nmax, Dmax, dout_max, prop);
  if (rc != GPC_OK) {  // e.g. out of memory at large N: give back what was allocated (the destroy tolerates null members)
    std::string keep = g_error;
    gpc_ctx_destroy(c);
    cudaGetLastError();
    g_error = keep;
    return rc;
  }
  *out = c;
  return GPC_OK;
}

static int ctx_create_impl(gpc_ctx* c, int device, int64_t Nmax, int Dmax, int dout_max, const cudaDeviceProp& prop) {
  c->device = device;
  c->Nmax = Nmax;
  c->Npmax = round_up(Nmax, TILE);
  c->Dmax = Dmax;
  c->dmax = dout_max;
  c->max_ctas = prop.multiProcessorCount * 2;
  {
    const char* mode = getenv("GPC_POTRF_MODE");
    c->use_winv = !(mode && std::string(mode) == "rec");
  }
  GPC_CUDA_CHECK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  size_t np = (size_t)c->Npmax;
  GPC_CUDA_CHECK(cudaMalloc(&c->X, np * Dmax * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&c->M, np * dout_max * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&c->alpha, np * dout_max * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&c->K, np * np * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&c->L, np * np * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&c->Dinv, np * TILE * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&c->scal, (SC_G + GPC_MAX_PARAMS) * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&c->info, sizeof(int)));
  GPC_CUDA_CHECK(cudaMalloc(&c->partial, (size_t)c->max_ctas * GPC_MAX_PARAMS * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&c->gXdev, np * Dmax * sizeof(double)));
  GPC_CUDA_CHECK(cudaMallocHost(&c->hres, (SC_G + GPC_MAX_PARAMS) * sizeof(double)));
  GPC_CUDA_CHECK(cudaMallocHost(&c->hinfo, sizeof(int)));
  GPC_CUDA_CHECK(cudaMemset(c->X, 0, np * Dmax * sizeof(double)));
  GPC_CUDA_CHECK(cudaMemset(c->M, 0, np * dout_max * sizeof(double)));
  GPC_CUDA_CHECK(cudaMemset(c->alpha, 0, np * dout_max * sizeof(double)));
  for (int i = 0; i < 6; i++) GPC_CUDA_CHECK(cudaEventCreate(&c->ev[i]));
  {
    int lo = 0, hi = 0;
    GPC_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = numerically lowest = highest priority
    GPC_CUDA_CHECK(cudaStreamCreateWithPriority(&c->chain_stream, cudaStreamNonBlocking, hi));
    GPC_CUDA_CHECK(cudaStreamCreateWithPriority(&c->bulk_stream, cudaStreamNonBlocking, lo));
    for (int i = 0; i < 6; i++) GPC_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_pipe[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) GPC_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_front[i], cudaEventDisableTiming));
  }
  if (!getenv("GPC_NO_FORK")) {
    c->fork = new Fork();
    c->fork->side.resize(16);
    // two events per node of the recursions, N/128 - 1 nodes each: the round-robin pool must not wrap inside one
    // evaluation (a re-recorded event would redirect a wait that has not been queued yet)
    c->fork->ev.resize((size_t)(8 * (c->Npmax / TILE) + 512));
    // the side streams carry pieces of the factorisation's own chain (panel copies, small products): chain priority
    int lo = 0, hi = 0;
    GPC_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (auto& st : c->fork->side) GPC_CUDA_CHECK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, hi));
    for (auto& e : c->fork->ev) GPC_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  return GPC_OK;
}

int gpc_ctx_destroy(gpc_ctx* c) {
  if (!c) return GPC_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaFree(c->X); cudaFree(c->M); cudaFree(c->alpha); cudaFree(c->K); cudaFree(c->L);
  cudaFree(c->Kinv); cudaFree(c->Winv); cudaFree(c->W); cudaFree(c->Dinv); cudaFree(c->scal); cudaFree(c->info);
  cudaFree(c->partial); cudaFree(c->gXdev); cudaFree(c->symm_part); cudaFree(c->zw); cudaFree(c->trmv_part); cudaFree(c->pipe_scratch); cudaFree(c->Xs); cudaFree(c->Kc); cudaFree(c->Kc2); cudaFree(c->tmp1); cudaFree(c->tmp2);
  cudaFreeHost(c->hres); cudaFreeHost(c->hinfo);
  gpc_ctx_set_profile(c, 0);
  if (c->fork) {
    for (auto st : c->fork->side)
      if (st) cudaStreamDestroy(st);
    for (auto e : c->fork->ev)
      if (e) cudaEventDestroy(e);
    delete c->fork;
  }
  for (int i = 0; i < 6; i++)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < 6; i++)
    if (c->ev_pipe[i]) cudaEventDestroy(c->ev_pipe[i]);
  for (int i = 0; i < 2; i++)
    if (c->ev_front[i]) cudaEventDestroy(c->ev_front[i]);
  if (c->chain_stream) cudaStreamDestroy(c->chain_stream);
  if (c->bulk_stream) cudaStreamDestroy(c->bulk_stream);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
  return GPC_OK;
}

int gpc_ctx_set_stream(gpc_ctx* c, void* s) {
  if (!c) return GPC_ERR_ARG;
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return GPC_OK;
}
void* gpc_ctx_get_stream(gpc_ctx* c) { return c ? (void*)c->stream : nullptr; }
int gpc_ctx_sync(gpc_ctx* c) {
  if (!c) return GPC_ERR_ARG;
  GPC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return GPC_OK;
}
int64_t gpc_ctx_launch_count(gpc_ctx* c) { return c ? c->launches : 0; }

static int upload(gpc_ctx* c, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols) {
  GPC_CUDA_CHECK(cudaMemcpy2DAsync(dst, ldd * sizeof(double), src, lds * sizeof(double), rows * sizeof(double), cols,
                                   cudaMemcpyHostToDevice, c->stream));
  return GPC_OK;
}
static int download(gpc_ctx* c, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows,
                    int64_t cols) {
  GPC_CUDA_CHECK(cudaMemcpy2DAsync(dst, ldd * sizeof(double), src, lds * sizeof(double), rows * sizeof(double), cols,
                                   cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return GPC_OK;
}

int gpc_set_X(gpc_ctx* c, const double* X, int64_t N, int D, int64_t ldx) {
  if (!c || !X || N < 1 || N > c->Nmax || D < 1 || D > c->Dmax || ldx < N) {
    set_error("gpc_set_X: bad arguments");
    return GPC_ERR_ARG;
  }
  GPC_CUDA_CHECK(cudaSetDevice(c->device));
  int64_t Np = round_up(N, TILE);
  if (c->haveX && (Np != c->Np || N != c->N)) {  // stale rows beyond the new N must read as zero
    GPC_CUDA_CHECK(cudaMemsetAsync(c->X, 0, (size_t)c->Npmax * c->Dmax * sizeof(double), c->stream));
    // m and alpha are laid out with the old leading dimension / row count: the caller has to set m again
    GPC_CUDA_CHECK(cudaMemsetAsync(c->M, 0, (size_t)c->Npmax * c->dmax * sizeof(double), c->stream));
    GPC_CUDA_CHECK(cudaMemsetAsync(c->alpha, 0, (size_t)c->Npmax * c->dmax * sizeof(double), c->stream));
    c->haveM = false;
  }
  c->N = N;
  c->Np = Np;
  c->D = D;
  GPC_CHECK(upload(c, c->X, Np, X, ldx, N, D));
  c->haveX = true;
  c->haveK = c->haveL = c->haveInv = c->haveAlpha = false;
  c->k_lazy = false;
  return GPC_OK;
}

int gpc_set_M(gpc_ctx* c, const double* M, int64_t N, int d, int64_t ldm) {
  if (!c || !M || !c->haveX || N != c->N || d < 1 || d > c->dmax || ldm < N) {
    set_error("gpc_set_M: bad arguments (set X first; N must match)");
    return GPC_ERR_ARG;
  }
  GPC_CUDA_CHECK(cudaSetDevice(c->device));
  GPC_CUDA_CHECK(cudaMemsetAsync(c->M, 0, (size_t)c->Npmax * c->dmax * sizeof(double), c->stream));
  GPC_CUDA_CHECK(cudaMemsetAsync(c->alpha, 0, (size_t)c->Npmax * c->dmax * sizeof(double), c->stream));
  c->d = d;
  GPC_CHECK(upload(c, c->M, c->Np, M, ldm, N, d));
  c->haveM = true;
  c->haveAlpha = false;
  return GPC_OK;
}

int gpc_set_Y(gpc_ctx* c, const double* Y, int64_t N, int d, int64_t ldy, const double* bias, const double* scale) {
  // m(:,j) = (y(:,j) - bias_j) / scale_j  (CGp::updateM, CGp.cpp:248-260): O(N d) host work, then one upload
  if (!c || !Y || N < 1 || d < 1 || ldy < N) {
    set_error("gpc_set_Y: bad arguments");
    return GPC_ERR_ARG;
  }
  std::vector<double> m((size_t)N * d);
  for (int j = 0; j < d; j++) {
    double b = bias ? bias[j] : 0.0, inv = 1.0 / (scale ? scale[j] : 1.0);
    for (int64_t i = 0; i < N; i++) m[i + (size_t)j * N] = (Y[i + (size_t)j * ldy] - b) * inv;
  }
  int rc = gpc_set_M(c, m.data(), N, d, N);
  if (rc == GPC_OK) rc = gpc_ctx_sync(c);  // m is a temporary
  return rc;
}

static int need(gpc_ctx* c, bool cond, const char* what) {
  if (!c) {
    set_error("null context");
    return GPC_ERR_ARG;
  }
  if (!cond) {
    set_error(std::string("state: ") + what);
    return GPC_ERR_STATE;
  }
  if (cudaSetDevice(c->device) != cudaSuccess) {
    set_error("cudaSetDevice failed");
    return GPC_ERR_CUDA;
  }
  return GPC_OK;
}

int gpc_kern_build(gpc_ctx* c, const gpc_kcomp* comps, int ncomp) {
  GPC_CHECK(need(c, c && c->haveX, "gpc_kern_build needs X"));
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, c->D, &ks));
  GPC_CHECK(launch_kbuild(ks, c->X, c->Np, c->N, c->Np, c->K, c->Np, c->stream, &c->launches));
  c->haveK = true;
  c->k_lazy = false;
  c->haveL = c->haveInv = c->haveAlpha = false;
  return GPC_OK;
}

// gpc_eval factors K in the buffer it was built in; whoever needs the K buffer afterwards gets it rebuilt here
static int ensure_K(gpc_ctx* c) {
  if (!c || c->haveK || !c->k_lazy || !c->haveX) return GPC_OK;
  GPC_CHECK(launch_kbuild(c->ks_last, c->X, c->Np, c->N, c->Np, c->K, c->Np, c->stream, &c->launches));
  c->haveK = true;
  c->k_lazy = false;
  return GPC_OK;
}

int gpc_add_diag(gpc_ctx* c, double jitter) {
  GPC_CHECK(ensure_K(c));
  GPC_CHECK(need(c, c && c->haveK, "gpc_add_diag needs K"));
  GPC_CHECK(launch_add_diag(c->K, c->Np, c->N, jitter, c->stream, &c->launches));
  c->haveL = c->haveInv = c->haveAlpha = false;
  return GPC_OK;
}

// in_place: the lower triangle of K is already in the L buffer (gpc_eval builds it there)
static int potrf_async(gpc_ctx* c, bool in_place = false) {
  GPC_CUDA_CHECK(cudaMemsetAsync(c->info, 0, sizeof(int), c->stream));
  GPC_CUDA_CHECK(cudaMemsetAsync(c->scal + SC_LOGDET, 0, sizeof(double), c->stream));
  if (!in_place) GPC_CHECK(launch_copy_lower(c->K, c->Np, c->L, c->Np, c->Np, c->stream, &c->launches));
  if (c->use_winv) {
    GPC_CHECK(ensure_inverse_buffers(c));
    if (c->winv_np != c->Np) {
      // the leaf writes only the lower part of W's diagonal blocks: zero the strict upper parts once per layout
      GPC_CUDA_CHECK(cudaMemsetAsync(c->Winv, 0, (size_t)c->Np * c->Np * sizeof(double), c->stream));
      c->winv_np = c->Np;
    }
    Dense d = dense_of(c);
    c->w_deferred = true;
    c->top_t_ready = nullptr;
    // only where the factorisation is latency-bound (idle SMs to fill); at N = 32768 two concurrent large products
    // just contend for L2
    if (c->inverse_follows && c->Np <= 16384) d.top_t_ready = &c->top_t_ready;
    return potrf_inv_rec(d, c->L, c->Np, c->Np, 0, d.Tpool, true, 0);
  }
  Dense d = dense_of(c);
  return potrf_rec(d, c->L, c->Np, c->Np, 0);
}

static int inverse_async(gpc_ctx* c) {
  Dense d = dense_of(c);
  if (c->use_winv) {
    d.top_t_ready = &c->top_t_ready;
    GPC_CHECK(inverse_from_W(d, c->L, c->Np, c->Np, c->Kinv, c->Np, c->w_deferred));
    c->w_deferred = false;
    return GPC_OK;
  }
  return potri_rec(d, c->L, c->Np, c->Np, c->Kinv, c->Np, 0);
}

int gpc_potrf(gpc_ctx* c, int* info, double* logdet) {
  GPC_CHECK(ensure_K(c));
  GPC_CHECK(need(c, c && c->haveK, "gpc_potrf needs K"));
  c->haveL = c->haveInv = c->haveAlpha = false;  // L is overwritten: valid again only if info == 0
  c->pipe.on = false;
  c->front.on = false;
  GPC_CHECK(potrf_async(c));
  GPC_CUDA_CHECK(cudaMemcpyAsync(c->hinfo, c->info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA_CHECK(cudaMemcpyAsync(c->hres, c->scal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (info) *info = *c->hinfo;
  if (logdet) *logdet = c->hres[SC_LOGDET];
  c->haveL = (*c->hinfo == 0);
  c->haveInv = c->haveAlpha = false;
  return *c->hinfo;
}

static int trace_K(gpc_ctx* c, double* tr) {
  // trace via a strided dot with ones is overkill: download the diagonal (N doubles)
  std::vector<double> diag((size_t)c->N);
  GPC_CUDA_CHECK(cudaMemcpy2DAsync(diag.data(), sizeof(double), c->K, (c->Np + 1) * sizeof(double), sizeof(double),
                                   c->N, cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  double t = 0.0;
  for (int64_t i = 0; i < c->N; i++) t += diag[i];
  *tr = t;
  return GPC_OK;
}

int gpc_jitchol(gpc_ctx* c, int max_tries, double* jitter_out, double* logdet) {
  GPC_CHECK(ensure_K(c));
  GPC_CHECK(need(c, c && c->haveK, "gpc_jitchol needs K"));
  if (max_tries <= 0) max_tries = 20;  // CMatrix.h:1060 default
  double jitter = 0.0;
  bool have_trace = false;
  int tries = 0;
  while (tries < max_tries) {
    int info = 0;
    int rc = gpc_potrf(c, &info, logdet);
    if (rc < 0) return rc;
    if (info == 0) {
      if (!have_trace) {
        double tr;
        GPC_CHECK(trace_K(c, &tr));
        jitter = 1e-6 * tr / (double)c->N;
      }
      if (jitter_out) *jitter_out = jitter;
      return GPC_OK;
    }
    if (!have_trace) {
      double tr;
      GPC_CHECK(trace_K(c, &tr));
      jitter = 1e-6 * tr / (double)c->N;  // CMatrix.cpp:775
      have_trace = true;
    }
    GPC_CHECK(gpc_add_diag(c, jitter));  // A.addDiag(jitter) mutates K (CMatrix.cpp:787)
    jitter *= 10.0;
    tries++;
    if (jitter > 10.0) {  // CMatrix.cpp:790-791
      if (jitter_out) *jitter_out = jitter;
      set_error("jitChol: matrix is non positive definite (jitter > 10)");
      return info;
    }
  }
  if (jitter_out) *jitter_out = jitter;
  set_error("jitChol: adding jitter failed after max tries");
  return 1;
}

int gpc_inverse(gpc_ctx* c) {
  GPC_CHECK(need(c, c && c->haveL, "gpc_inverse needs a successful gpc_potrf"));
  GPC_CHECK(ensure_inverse_buffers(c));
  GPC_CHECK(inverse_async(c));
  c->haveInv = true;
  c->kinv_full = true;
  return GPC_OK;
}

static int alpha_from_inverse_async(gpc_ctx* c) {
  if (!c->kinv_full) {  // the symmetric product reads whole rows
    GPC_CHECK(launch_mirror_lower(c->Kinv, c->Np, c->Np, c->stream, &c->launches));
    c->kinv_full = true;
  }
  GPC_CUDA_CHECK(cudaMemsetAsync(c->scal + SC_QUAD, 0, sizeof(double), c->stream));
  // scratch for the column-chunk partial sums: the inverse workspace W is idle once K^-1 is complete
  GPC_CHECK(launch_symm_small(c->Kinv, c->Np, c->M, c->Np, c->alpha, c->Np, c->N, c->d, c->symm_part, c->stream,
                              &c->launches));
  return launch_dot(c->M, c->alpha, c->Np * c->d, c->scal + SC_QUAD, c->stream, &c->launches);
}

// alpha = W'(W m) and quad = m' alpha on stream s, straight from the (complete) triangular inverse: two HBM-bound
// triangular matrix-vector products that do not wait for K^-1 = W'W
static int alpha_from_W_async(gpc_ctx* c, cudaStream_t s) {
  GPC_CUDA_CHECK(cudaMemsetAsync(c->scal + SC_QUAD, 0, sizeof(double), s));
  GPC_CHECK(launch_trmv_lower(c->Winv, c->Np, false, c->M, c->Np, c->zw, c->Np, c->Np, c->d, c->trmv_part, s, &c->launches));
  GPC_CHECK(launch_trmv_lower(c->Winv, c->Np, true, c->zw, c->Np, c->alpha, c->Np, c->Np, c->d, c->trmv_part, s, &c->launches));
  return launch_dot(c->M, c->alpha, c->Np * c->d, c->scal + SC_QUAD, s, &c->launches);
}

int gpc_alpha_from_inverse(gpc_ctx* c, double* quad) {
  GPC_CHECK(need(c, c && c->haveInv && c->haveM, "gpc_alpha_from_inverse needs K^-1 and m"));
  GPC_CHECK(alpha_from_inverse_async(c));
  GPC_CUDA_CHECK(cudaMemcpyAsync(c->hres, c->scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (quad) *quad = c->hres[SC_QUAD];
  c->haveAlpha = true;
  return GPC_OK;
}

static int ensure_tmp(gpc_ctx* c, int64_t elems) {
  if (c->tmp_cap >= elems) return GPC_OK;
  cudaFree(c->tmp1);
  cudaFree(c->tmp2);
  c->tmp1 = c->tmp2 = nullptr;
  c->tmp_cap = 0;  // stays 0 if the second allocation fails: the pair is re-allocated by the next call
  GPC_CUDA_CHECK(cudaMalloc(&c->tmp1, (size_t)elems * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&c->tmp2, (size_t)elems * sizeof(double)));
  c->tmp_cap = elems;
  return GPC_OK;
}

// alpha = L^-T L^-1 m through the triangular factor: work on the transposed right-hand side (dp x Np),
// dp = TILE-padded output count, so that both solves are right-sided: Z L' = M', then A L = Z.
int gpc_solve_alpha(gpc_ctx* c, double* quad) {
  GPC_CHECK(need(c, c && c->haveL && c->haveM, "gpc_solve_alpha needs L and m"));
  int64_t dp = round_up(c->d, TILE);
  GPC_CHECK(ensure_tmp(c, dp * c->Np));
  Dense d = dense_of(c);
  GPC_CUDA_CHECK(cudaMemsetAsync(c->tmp1, 0, (size_t)dp * c->Np * sizeof(double), c->stream));
  GPC_CHECK(launch_transpose(c->M, c->Np, c->tmp1, dp, c->Np, c->d, c->stream, &c->launches));
  GPC_CHECK(trsm_rlt(d, c->tmp1, dp, dp, c->L, c->Np, c->Np, 0));
  GPC_CHECK(trsm_rln(d, c->tmp1, dp, dp, c->L, c->Np, c->Np, 0));
  GPC_CHECK(launch_transpose(c->tmp1, dp, c->alpha, c->Np, c->d, c->Np, c->stream, &c->launches));
  GPC_CUDA_CHECK(cudaMemsetAsync(c->scal + SC_QUAD, 0, sizeof(double), c->stream));
  GPC_CHECK(launch_dot(c->M, c->alpha, c->Np * c->d, c->scal + SC_QUAD, c->stream, &c->launches));
  GPC_CUDA_CHECK(cudaMemcpyAsync(c->hres, c->scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (quad) *quad = c->hres[SC_QUAD];
  c->haveAlpha = true;
  return GPC_OK;
}

static int grad_async(gpc_ctx* c, const KSpec& ks, bool wantX) {
  if (wantX) GPC_CUDA_CHECK(cudaMemsetAsync(c->gXdev, 0, (size_t)c->Np * c->D * sizeof(double), c->stream));
  return launch_grad(ks, c->X, c->Np, c->N, c->Np, c->Kinv, c->Np, c->alpha, c->Np, c->d, 0, c->partial, c->max_ctas,
                     c->scal + SC_G, wantX ? c->gXdev : nullptr, c->Np, c->stream, &c->launches);
}

int gpc_grad(gpc_ctx* c, const gpc_kcomp* comps, int ncomp, double* gparams, double* gX) {
  GPC_CHECK(need(c, c && c->haveInv && c->haveAlpha, "gpc_grad needs K^-1 and alpha"));
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, c->D, &ks));
  GPC_CHECK(grad_async(c, ks, gX != nullptr));
  GPC_CUDA_CHECK(cudaMemcpyAsync(c->hres + SC_G, c->scal + SC_G, ks.nparams * sizeof(double), cudaMemcpyDeviceToHost,
                                 c->stream));
  if (gX) GPC_CHECK(download(c, gX, c->N, c->gXdev, c->Np, c->N, c->D));
  GPC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < ks.nparams; i++) gparams[i] = c->hres[SC_G + i];
  return GPC_OK;
}

int gpc_kern_grad(gpc_ctx* c, const gpc_kcomp* comps, int ncomp, const double* covGrad, int64_t ldc, double* gparams,
                  double* gX) {
  GPC_CHECK(need(c, c && c->haveX, "gpc_kern_grad needs X"));
  if (!covGrad || ldc < c->N) {
    set_error("gpc_kern_grad: bad covGrad");
    return GPC_ERR_ARG;
  }
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, c->D, &ks));
  GPC_CHECK(ensure_inverse_buffers(c));
  // stage covGrad in the K^-1 buffer (invalidates it)
  c->haveInv = false;
  GPC_CHECK(upload(c, c->Kinv, c->Np, covGrad, ldc, c->N, c->N));
  if (gX) GPC_CUDA_CHECK(cudaMemsetAsync(c->gXdev, 0, (size_t)c->Np * c->D * sizeof(double), c->stream));
  GPC_CHECK(launch_grad(ks, c->X, c->Np, c->N, c->Np, c->Kinv, c->Np, nullptr, 0, 0, 1, c->partial, c->max_ctas,
                        c->scal + SC_G, gX ? c->gXdev : nullptr, c->Np, c->stream, &c->launches));
  GPC_CUDA_CHECK(cudaMemcpyAsync(c->hres + SC_G, c->scal + SC_G, ks.nparams * sizeof(double), cudaMemcpyDeviceToHost,
                                 c->stream));
  if (gX) GPC_CHECK(download(c, gX, c->N, c->gXdev, c->Np, c->N, c->D));
  GPC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < ks.nparams; i++) gparams[i] = c->hres[SC_G + i];
  return GPC_OK;
}

static int ensure_cross(gpc_ctx* c, int64_t Nsp);

// sum_ij covGrad2[i,j] dk(X_i, X2_j)/dtheta (natural parameters) and, optionally, gX[i,:] = sum_j covGrad2[i,j] dk(X_i,X2_j)/dX_i:
// CKern::getGradParams(g, X, X2, covGrad) (CKern.h:199-213; rbf CKern.cpp:1204-1241 ...) and the covGrad-weighted
// CKern::getGradX (CKern.h:68-74), computeElement semantics (white contributes nothing).
int gpc_kern_grad_cross(gpc_ctx* c, const gpc_kcomp* comps, int ncomp, const double* X2, int64_t N2, int64_t ldx2,
                        const double* covGrad2, int64_t ldc, double* gparams, double* gX) {
  GPC_CHECK(need(c, c && c->haveX, "gpc_kern_grad_cross needs X"));
  if (!X2 || !covGrad2 || !gparams || N2 < 1 || ldx2 < N2 || ldc < c->N) {
    set_error("gpc_kern_grad_cross: bad arguments");
    return GPC_ERR_ARG;
  }
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, c->D, &ks));
  const int64_t N2p = round_up(N2, TILE);
  GPC_CHECK(ensure_cross(c, N2p));
  GPC_CUDA_CHECK(cudaMemsetAsync(c->Xs, 0, (size_t)N2p * c->D * sizeof(double), c->stream));
  GPC_CHECK(upload(c, c->Xs, N2p, X2, ldx2, N2, c->D));
  GPC_CHECK(upload(c, c->Kc, c->Np, covGrad2, ldc, c->N, N2));
  if (gX) GPC_CUDA_CHECK(cudaMemsetAsync(c->gXdev, 0, (size_t)c->Np * c->D * sizeof(double), c->stream));
  GradCross cx{1, c->Xs, N2p, N2};
  GPC_CHECK(launch_grad(ks, c->X, c->Np, c->N, c->Np, c->Kc, c->Np, nullptr, 0, 0, 1, c->partial, c->max_ctas,
                        c->scal + SC_G, gX ? c->gXdev : nullptr, c->Np, c->stream, &c->launches, -1, 0, nullptr, &cx));
  GPC_CUDA_CHECK(cudaMemcpyAsync(c->hres + SC_G, c->scal + SC_G, ks.nparams * sizeof(double), cudaMemcpyDeviceToHost,
                                 c->stream));
  if (gX) GPC_CHECK(download(c, gX, c->N, c->gXdev, c->Np, c->N, c->D));
  GPC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < ks.nparams; i++) gparams[i] = c->hres[SC_G + i];
  return GPC_OK;
}

// V = B L^-T (the right-sided solve V L' = B) through W = L^-1, out of place: one triangular product when W is
// complete; while the top-level W21 is still deferred the top level is done by block substitution instead
// (V1 = B1 W11', B2 -= V1 L21', V2 = B2 W22'), which needs only the diagonal halves of W.  B is overwritten.
static int solve_rlt_via_W(gpc_ctx* c, const Dense& d, double* B, int64_t ldb, int64_t m, double* V, int64_t ldv) {
  const int64_t n = c->Np;
  if (!c->w_deferred || n <= TILE) {
    GemmCall g{B, d.Winv, V, ldb, d.ldw, ldv, m, n, n, 1.0, 0.0, false, false, false};
    g.b_tri = -1;
    return launch_gemm(g, d.s, d.launches);
  }
  const int64_t n1 = (n / TILE / 2) * TILE, n2 = n - n1;
  {
    GemmCall g{B, d.Winv, V, ldb, d.ldw, ldv, m, n1, n1, 1.0, 0.0, false, false, false};
    g.b_tri = -1;
    GPC_CHECK(launch_gemm(g, d.s, d.launches));
  }
  {  // B2 -= V1 L21'
    GemmCall g{V, c->L + n1, B + n1 * ldb, ldv, c->Np, ldb, m, n2, n1, -1.0, 1.0, false, false, false};
    GPC_CHECK(launch_gemm(g, d.s, d.launches));
  }
  GemmCall g{B + n1 * ldb, d.Winv + n1 + n1 * d.ldw, V + n1 * ldv, ldb, d.ldw, ldv, m, n2, n2, 1.0, 0.0, false, false, false};
  g.b_tri = -1;
  return launch_gemm(g, d.s, d.launches);
}

static int ensure_cross(gpc_ctx* c, int64_t Nsp) {
  if (c->Xs_cap < Nsp * c->Dmax) {
    cudaFree(c->Xs);
    c->Xs = nullptr;
    c->Xs_cap = 0;
    GPC_CUDA_CHECK(cudaMalloc(&c->Xs, (size_t)Nsp * c->Dmax * sizeof(double)));
    c->Xs_cap = Nsp * c->Dmax;
  }
  if (c->Kc_cap < Nsp * c->Np) {
    cudaFree(c->Kc);
    cudaFree(c->Kc2);
    c->Kc = c->Kc2 = nullptr;
    c->Kc_cap = 0;
    GPC_CUDA_CHECK(cudaMalloc(&c->Kc, (size_t)Nsp * c->Np * sizeof(double)));
    if (c->use_winv) GPC_CUDA_CHECK(cudaMalloc(&c->Kc2, (size_t)Nsp * c->Np * sizeof(double)));
    c->Kc_cap = Nsp * c->Np;
  }
  return GPC_OK;
}

int gpc_kern_cross(gpc_ctx* c, const gpc_kcomp* comps, int ncomp, const double* Xs, int64_t Ns, int64_t ldxs,
                   double* Ks, int64_t ldk) {
  GPC_CHECK(need(c, c && c->haveX, "gpc_kern_cross needs X"));
  if (!Xs || !Ks || Ns < 1 || ldxs < Ns || ldk < c->N) {
    set_error("gpc_kern_cross: bad arguments");
    return GPC_ERR_ARG;
  }
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, c->D, &ks));
  int64_t Nsp = round_up(Ns, TILE);
  GPC_CHECK(ensure_cross(c, Nsp));
  GPC_CHECK(upload(c, c->Xs, Nsp, Xs, ldxs, Ns, c->D));
  GPC_CHECK(launch_kcross(ks, c->X, c->Np, c->N, c->Np, c->Xs, Nsp, Ns, Nsp, c->Kc, c->Np, c->stream, &c->launches));
  return download(c, Ks, ldk, c->Kc, c->Np, c->N, Ns);
}

int gpc_kern_diag(gpc_ctx* c, const gpc_kcomp* comps, int ncomp, const double* Xs, int64_t Ns, int64_t ldxs,
                  double* kdiag) {
  GPC_CHECK(need(c, c != nullptr, "ctx"));
  if (!Xs || !kdiag || Ns < 1 || ldxs < Ns || !c->haveX) {
    set_error("gpc_kern_diag: bad arguments");
    return GPC_ERR_ARG;
  }
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, c->D, &ks));
  int64_t Nsp = round_up(Ns, TILE);
  GPC_CHECK(ensure_cross(c, Nsp));
  GPC_CHECK(ensure_tmp(c, Nsp));
  GPC_CHECK(upload(c, c->Xs, Nsp, Xs, ldxs, Ns, c->D));
  GPC_CHECK(launch_kdiag(ks, c->Xs, Nsp, Ns, c->tmp1, c->stream, &c->launches));
  return download(c, kdiag, Ns, c->tmp1, Nsp, Ns, 1);
}

int gpc_posterior(gpc_ctx* c, const gpc_kcomp* comps, int ncomp, const double* Xs, int64_t Ns, int64_t ldxs,
                  double* mu, double* var) {
  GPC_CHECK(need(c, c && c->haveL && c->haveAlpha, "gpc_posterior needs L and alpha"));
  if (!Xs || !mu || Ns < 1 || ldxs < Ns) {
    set_error("gpc_posterior: bad arguments");
    return GPC_ERR_ARG;
  }
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, c->D, &ks));
  int64_t Nsp = round_up(Ns, TILE);
  GPC_CHECK(ensure_cross(c, Nsp));
  GPC_CHECK(ensure_tmp(c, Nsp * (c->d > 1 ? c->d : 1) + Nsp));
  GPC_CHECK(upload(c, c->Xs, Nsp, Xs, ldxs, Ns, c->D));
  // Kc = K(Xs, X): Nsp x Np, test points along rows so that the solve is right-sided: V L' = Kc
  GPC_CHECK(launch_kcross(ks, c->Xs, Nsp, Ns, Nsp, c->X, c->Np, c->N, c->Np, c->Kc, Nsp, c->stream, &c->launches));
  // mu = Kc alpha  (CGp::_posteriorMean, CGp.cpp:548-560)
  GPC_CHECK(launch_gemv_rows(c->Kc, Nsp, Ns, c->N, c->alpha, c->Np, c->d, c->tmp1, Nsp, c->stream, &c->launches));
  GPC_CHECK(download(c, mu, Ns, c->tmp1, Nsp, Ns, c->d));
  if (var) {
    // var = k(x*,x*) - |L^-1 k*|^2  (CGp::_posteriorVar FTC, CGp.cpp:600-613)
    Dense d = dense_of(c);
    const double* Vm = c->Kc;
    if (c->use_winv && c->Kc2) {  // V = Kc L^-T as products with W = L^-1 (large GEMMs on the tensor-core engine)
      GPC_CHECK(solve_rlt_via_W(c, d, c->Kc, Nsp, Nsp, c->Kc2, Nsp));
      Vm = c->Kc2;
    } else {
      GPC_CHECK(trsm_rlt(d, c->Kc, Nsp, Nsp, c->L, c->Np, c->Np, 0));
    }
    GPC_CHECK(launch_kdiag(ks, c->Xs, Nsp, Ns, c->tmp2, c->stream, &c->launches));
    GPC_CHECK(launch_row_sqnorm_sub(Vm, Nsp, Ns, c->N, c->tmp2, c->tmp1, c->stream, &c->launches));
    std::vector<double> v((size_t)Ns);
    GPC_CHECK(download(c, v.data(), Ns, c->tmp1, Nsp, Ns, 1));
    for (int j = 0; j < c->d; j++)  // same variance for every output (CGp.cpp:608-611)
      memcpy(var + (size_t)j * Ns, v.data(), sizeof(double) * Ns);
  }
  return GPC_OK;
}

static int trace_phase(gpc_ctx* c, const char* what) {
  static int on = -1;
  if (on < 0) on = getenv("GPC_TRACE") ? 1 : 0;
  if (!on) return GPC_OK;
  cudaError_t e = cudaStreamSynchronize(c->stream);
  fprintf(stderr, "[gpc trace] %s: %s (launches so far %lld)\n", what, cudaGetErrorString(e), (long long)c->launches);
  fflush(stderr);
  if (e != cudaSuccess) {
    set_error(std::string("trace: ") + what + ": " + cudaGetErrorString(e));
    return GPC_ERR_CUDA;
  }
  return GPC_OK;
}

int gpc_eval(gpc_ctx* c, const gpc_kcomp* comps, int ncomp, int flags, double* out, double* gparams, double* gX) {
  GPC_CHECK(need(c, c && c->haveX && c->haveM, "gpc_eval needs X and m"));
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, c->D, &ks));
  GPC_CHECK(ensure_inverse_buffers(c));
  bool wantX = (flags & 1) && gX;
  Dense d = dense_of(c);
  cudaStream_t s = c->stream;
  // L, W, K^-1 and alpha are overwritten from here on: they only become valid again on the success path
  c->haveL = c->haveInv = c->haveAlpha = false;
  const auto host_t0 = std::chrono::steady_clock::now();
  GPC_CUDA_CHECK(cudaEventRecord(c->ev[0], s));
  // K is built straight into the buffer that is factored in place (its lower triangle is all the factorisation reads);
  // the K buffer itself is rebuilt only if somebody asks for it (ensure_K) or the jitter schedule needs to mutate it
  GPC_CHECK(launch_kbuild(ks, c->X, c->Np, c->N, c->Np, c->L, c->Np, s, &c->launches));
  c->haveK = false;
  c->k_lazy = true;
  c->ks_last = ks;
  bool in_place = true;
  GPC_CUDA_CHECK(cudaEventRecord(c->ev[1], s));
  GPC_CHECK(trace_phase(c, "kbuild"));
  double jitter_used = 0.0;
  double jitter = 0.0;
  bool have_trace = false;
  for (int tries = 0;; tries++) {
    if (c->prof) {
      c->prof->used = 0;
      c->prof->flops.clear();
      c->prof->recs.clear();
    }
    // the factorisation on the high-priority chain stream; the row-block pipeline of the top level (TopPipe) on the
    // low-priority bulk stream when its scratch exists (N <= 16384) and the second half is not a single diagonal block
    static const int chain_prio = getenv("GPC_CHAIN_PRIO") ? atoi(getenv("GPC_CHAIN_PRIO")) : 1;
    // (where the factorisation is latency-bound: beyond N = 16384 the large products own the GPU anyway)
    const bool hp = chain_prio && c->use_winv && c->fork && !c->prof && c->chain_stream && c->Np <= 16384;
    const bool pipe = hp && c->pipe_scratch && c->bulk_stream && c->Np >= 1024;
    c->pipe.on = pipe;
    c->pipe.y_queued = false;
    c->pipe.t_ready = nullptr;
    c->pipe.Kinv = c->Kinv;
    c->pipe.ldo = c->Np;
    c->pipe.bulk = c->bulk_stream;
    c->pipe.e_w11 = c->ev_pipe[2];
    c->pipe.e_half = c->ev_pipe[3];
    c->pipe.e_bulk = c->ev_pipe[4];
    static const int top_front = getenv("GPC_TOP_FRONT") ? atoi(getenv("GPC_TOP_FRONT")) : 1;
    c->front.on = hp && top_front && c->bulk_stream;
    c->front.queued = false;
    c->front.n1 = 0;
    c->front.e_a = c->ev_front[0];
    c->front.e_x2 = c->ev_front[1];
    if (hp) {
      GPC_CUDA_CHECK(cudaEventRecord(c->ev_pipe[0], s));
      GPC_CUDA_CHECK(cudaStreamWaitEvent(c->chain_stream, c->ev_pipe[0], 0));
      c->stream = c->chain_stream;
    }
    c->inverse_follows = true;
    int prc = potrf_async(c, in_place);
    c->inverse_follows = false;
    if (hp) {
      c->stream = s;
      if (prc == GPC_OK) {
        GPC_CUDA_CHECK(cudaEventRecord(c->ev_pipe[1], c->chain_stream));
        GPC_CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_pipe[1], 0));
      }
    }
    GPC_CHECK(prc);
    GPC_CUDA_CHECK(cudaEventRecord(c->ev[2], s));
    GPC_CHECK(trace_phase(c, "potrf"));
    // optimistic: queue the rest before looking at info (a failed factorisation is rare and just redone)
    if (c->use_winv) {
      Dense dd = dense_of(c);
      dd.top_t_ready = &c->top_t_ready;
      GPC_CHECK(complete_W(dd, c->L, c->Np, c->Np, c->w_deferred));
      c->w_deferred = false;
      // alpha = W'(W m) needs W only: two HBM-bound products on a side stream, hidden behind the tensor-bound W'W
      cudaEvent_t e_alpha = nullptr;
      if (c->fork && !c->prof) {
        cudaEvent_t e_w = c->fork->event();
        e_alpha = c->fork->event();
        // low priority: the two products must not take SMs from the W'W product they hide behind
        cudaStream_t side = c->bulk_stream ? c->bulk_stream : c->fork->stream();
        GPC_CUDA_CHECK(cudaEventRecord(e_w, s));
        GPC_CUDA_CHECK(cudaStreamWaitEvent(side, e_w, 0));
        GPC_CHECK(alpha_from_W_async(c, side));
        GPC_CUDA_CHECK(cudaEventRecord(e_alpha, side));
      }
      if (pipe) {  // whatever the pipeline queued on the bulk stream writes K^-1: join it before the rest is added
        GPC_CUDA_CHECK(cudaEventRecord(c->ev_pipe[5], c->bulk_stream));
        GPC_CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_pipe[5], 0));
      }
      GPC_CHECK(kinv_from_W(dd, c->Np, c->Kinv, c->Np, false));  // lower triangle only: all the gradient pass reads
      c->pipe.on = false;
      c->kinv_full = false;
      GPC_CUDA_CHECK(cudaEventRecord(c->ev[3], s));
      GPC_CHECK(trace_phase(c, "inverse"));
      if (e_alpha) GPC_CUDA_CHECK(cudaStreamWaitEvent(s, e_alpha, 0));
      else GPC_CHECK(alpha_from_W_async(c, s));
    } else {
      GPC_CHECK(inverse_async(c));
      c->kinv_full = true;
      GPC_CUDA_CHECK(cudaEventRecord(c->ev[3], s));
      GPC_CHECK(trace_phase(c, "inverse"));
      GPC_CHECK(alpha_from_inverse_async(c));
    }
    GPC_CUDA_CHECK(cudaEventRecord(c->ev[4], s));
    GPC_CHECK(trace_phase(c, "alpha"));
    GPC_CHECK(grad_async(c, ks, wantX));
    GPC_CUDA_CHECK(cudaEventRecord(c->ev[5], s));
    GPC_CHECK(trace_phase(c, "grad"));
    GPC_CUDA_CHECK(cudaMemcpyAsync(c->hinfo, c->info, sizeof(int), cudaMemcpyDeviceToHost, s));
    GPC_CUDA_CHECK(cudaMemcpyAsync(c->hres, c->scal, (SC_G + ks.nparams) * sizeof(double), cudaMemcpyDeviceToHost, s));
    c->enqueue_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    GPC_CUDA_CHECK(cudaStreamSynchronize(s));
    if (*c->hinfo == 0) break;
    // jitChol schedule (CMatrix.cpp:767-804): from here on K lives in its own buffer and is mutated there
    if (in_place) {
      GPC_CHECK(launch_kbuild(ks, c->X, c->Np, c->N, c->Np, c->K, c->Np, s, &c->launches));
      c->haveK = true;
      c->k_lazy = false;
      in_place = false;
    }
    if (!have_trace) {
      double tr;
      GPC_CHECK(trace_K(c, &tr));
      jitter = 1e-6 * tr / (double)c->N;
      have_trace = true;
    }
    GPC_CHECK(launch_add_diag(c->K, c->Np, c->N, jitter, s, &c->launches));
    jitter_used += jitter;
    jitter *= 10.0;
    if (jitter > 10.0 || tries + 1 >= 20) {
      set_error("gpc_eval: kernel matrix is non positive definite after jitter retries");
      return *c->hinfo;
    }
  }
  c->haveL = c->haveInv = c->haveAlpha = true;
  if (out) {
    out[0] = c->hres[SC_LOGDET];
    out[1] = c->hres[SC_QUAD];
    out[2] = jitter_used;
  }
  if (gparams)
    for (int i = 0; i < ks.nparams; i++) gparams[i] = c->hres[SC_G + i];
  if (wantX) GPC_CHECK(download(c, gX, c->N, c->gXdev, c->Np, c->N, c->D));
  if (c->prof) {
    c->prof_ms = 0.0;
    c->prof_flops = 0.0;
    c->prof_count = (int64_t)(c->prof->used / 2);
    for (int q = 0; q < 8; q++) c->prof_split[q] = 0.0;
    const char* dump = getenv("GPC_PROF_DUMP");
    FILE* df = dump ? fopen(dump, "w") : nullptr;
    if (df) fprintf(df, "m,n,k,lower,engine,ms,flops\n");
    const double S = (double)oz_slices();
    for (size_t i = 0; i + 1 < c->prof->used; i += 2) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, c->prof->ev[i], c->prof->ev[i + 1]);
      c->prof_ms += ms;
      c->prof_flops += c->prof->flops[i / 2];
      const GemmProf::Rec& r = c->prof->recs[i / 2];
      const int o = r.ozaki ? 4 : 0;  // [ms, flops, count, int8 ops] per engine
      c->prof_split[o + 0] += ms;
      c->prof_split[o + 1] += c->prof->flops[i / 2];
      c->prof_split[o + 2] += 1.0;
      if (r.ozaki) c->prof_split[o + 3] += c->prof->flops[i / 2] * S * (S + 1.0) / 2.0;
      if (df) fprintf(df, "%lld,%lld,%lld,%d,%s,%.6f,%.0f\n", (long long)r.m, (long long)r.n, (long long)r.k, r.lower,
                      r.ozaki ? "ozaki" : "dmma", ms, c->prof->flops[i / 2]);
    }
    if (df) fclose(df);
    c->prof->used = 0;
    c->prof->flops.clear();
    c->prof->recs.clear();
  }
  for (int i = 0; i < 5; i++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]);
    c->last_ms[i] = ms;
  }
  float tot = 0.f;
  cudaEventElapsedTime(&tot, c->ev[0], c->ev[5]);
  c->last_ms[5] = tot;
  return GPC_OK;
}

int gpc_ctx_set_profile(gpc_ctx* c, int on) {
  if (!c) return GPC_ERR_ARG;
  if (on && !c->prof) c->prof = new GemmProf();
  if (!on && c->prof) {
    for (cudaEvent_t e : c->prof->ev) cudaEventDestroy(e);
    delete c->prof;
    c->prof = nullptr;
  }
  return GPC_OK;
}
int gpc_last_gemm_profile(gpc_ctx* c, double* total_ms, int64_t* count, double* flops) {
  if (!c) return GPC_ERR_ARG;
  if (total_ms) *total_ms = c->prof_ms;
  if (count) *count = c->prof_count;
  if (flops) *flops = c->prof_flops;
  return GPC_OK;
}

int gpc_last_gemm_profile_split(gpc_ctx* c, double* out8) {
  if (!c || !out8) return GPC_ERR_ARG;
  for (int q = 0; q < 8; q++) out8[q] = c->prof_split[q];
  return GPC_OK;
}

int gpc_last_timings(gpc_ctx* c, double* ms6) {
  if (!c || !ms6) return GPC_ERR_ARG;
  for (int i = 0; i < 6; i++) ms6[i] = c->last_ms[i];
  return GPC_OK;
}
int gpc_ctx_dims(gpc_ctx* c, int64_t* N, int* D, int* d) {
  if (!c || !c->haveX || !c->haveM) {
    set_error("gpc_ctx_dims: X and m have not been set");
    return GPC_ERR_STATE;
  }
  if (N) *N = c->N;
  if (D) *D = c->D;
  if (d) *d = c->d;
  return GPC_OK;
}
int gpc_last_enqueue_ms(gpc_ctx* c, double* ms) {
  if (!c || !ms) return GPC_ERR_ARG;
  *ms = c->enqueue_ms;
  return GPC_OK;
}

int gpc_download(gpc_ctx* c, int which, double* dst, int64_t ld) {
  GPC_CHECK(need(c, c && dst, "gpc_download"));
  int64_t N = c->N;
  switch (which) {
    case GPC_MAT_K:
      GPC_CHECK(ensure_K(c));
      GPC_CHECK(need(c, c->haveK, "K not built"));
      GPC_CHECK(download(c, dst, ld, c->K, c->Np, N, N));
      for (int64_t j = 0; j < N; j++)
        for (int64_t i = 0; i < j; i++) dst[i + j * ld] = dst[j + i * ld];  // device keeps the lower triangle
      return GPC_OK;
    case GPC_MAT_L:
      GPC_CHECK(need(c, c->haveL, "L not available"));
      GPC_CHECK(download(c, dst, ld, c->L, c->Np, N, N));
      for (int64_t j = 0; j < N; j++)
        for (int64_t i = 0; i < j; i++) dst[i + j * ld] = 0.0;
      return GPC_OK;
    case GPC_MAT_KINV:
      GPC_CHECK(need(c, c->haveInv, "K^-1 not available"));
      GPC_CHECK(download(c, dst, ld, c->Kinv, c->Np, N, N));
      if (!c->kinv_full)  // gpc_eval writes the lower triangle only
        for (int64_t j = 0; j < N; j++)
          for (int64_t i = 0; i < j; i++) dst[i + j * ld] = dst[j + i * ld];
      return GPC_OK;
    case GPC_MAT_ALPHA:
      GPC_CHECK(need(c, c->haveAlpha, "alpha not available"));
      return download(c, dst, ld, c->alpha, c->Np, N, c->d);
    case GPC_MAT_M:
      GPC_CHECK(need(c, c->haveM, "m not set"));
      return download(c, dst, ld, c->M, c->Np, N, c->d);
  }
  set_error("gpc_download: unknown matrix");
  return GPC_ERR_ARG;
}

}  // extern "C"
