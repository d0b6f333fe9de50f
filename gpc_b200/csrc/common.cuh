// common.cuh -- shared declarations for libgpc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/gpc_b200.h"

namespace gpc {

constexpr int TILE = 128;  // every device matrix is padded to a multiple of TILE in both dimensions
inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

void set_error(const std::string& s);
// the device current to the calling thread: function attributes (opt-in shared memory sizes) are per device, so the
// "already configured" flags of the launch helpers are kept per device
inline int cur_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}
// GPC_TRACE=1: synchronise after every kernel launch and log its name (bring-up / hang localisation)
int trace_sync(const char* what, cudaStream_t s);
#define GPC_CUDA_CHECK(expr)                                                                        \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      gpc::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" +    \
                     std::to_string(__LINE__));                                                     \
      return GPC_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

// ---- SM partition (smpart.cu): a small "chain" set and a large "bulk" set of SMs (green contexts) --------------
struct SmPartition;
SmPartition* smpart_create(int device, int chain_sms, int dflt);  // null: switched off (GPC_SM_PARTITION=0) / unsupported
void smpart_destroy(SmPartition* p);
int smpart_sms(const SmPartition* p, bool chain);
int smpart_stream(SmPartition* p, bool chain, int priority, cudaStream_t* out);

// ---- kernel specification passed by value to the K-build / gradient kernels --------------------------
struct KSpec {
  int ncomp;
  int D;
  int nparams;                    // total
  int need_r2, need_dot;          // which pair quantities any component consumes
  int type[GPC_MAX_COMPONENTS];
  int poff[GPC_MAX_COMPONENTS];   // offset of the component's first parameter in p[]
  double degree[GPC_MAX_COMPONENTS];
  double p[GPC_MAX_PARAMS];
};
int make_kspec(const gpc_kcomp* comps, int ncomp, int D, KSpec* out);

// ---- dense layer (dense.cu) ---------------------------------------------------------------------------
// C(m x n) = alpha * opA(A)(m x k) * opB(B)(k x n) + beta * C.   All dims multiples of TILE (k of 16).
//   a_kc: A is stored with k contiguous, i.e. opA(A)(i,kk) = A[kk + i*lda]   ('T' in BLAS terms)
//         otherwise                           opA(A)(i,kk) = A[i + kk*lda]   ('N')
//   b_kc: B is stored with k contiguous, i.e. opB(B)(kk,j) = B[kk + j*ldb]   ('N' in BLAS terms)
//         otherwise                           opB(B)(kk,j) = B[j + kk*ldb]   ('T')
//   lower: only output tiles that intersect the lower triangle are computed (m == n required)
struct GemmCall {
  const double* A;
  const double* B;
  double* C;
  int64_t lda, ldb, ldc;
  int64_t m, n, k;
  double alpha, beta;
  bool a_kc, b_kc, lower;
  bool ktri = false;  // A operand is zero for kk < i (e.g. rows of L^-T): each row tile starts its k loop at its row
  // triangular operands (tile-granular k-range skipping; the skipped parts of the operands are never read):
  //   a_tri = +1: opA(A)(i, kk) == 0 for kk < i (same as ktri);  a_tri = -1: == 0 for kk > i
  //   b_tri = +1: opB(B)(kk, j) == 0 for kk < j;                  b_tri = -1: == 0 for kk > j
  int a_tri = 0, b_tri = 0;
  // > 0: at most this many SMs at a time for a product on the tensor-core engine (it is launched wave by wave, each
  // launch holding <= sm_limit resident CTAs): bulk products that run NEXT TO the serial chain of the factorisation
  // leave the other SMs free for it -- a tile owns its SM for ~50 us, and stream priorities alone do not help a chain
  // of small kernels that needs a free SM every few microseconds
  int sm_limit = 0;
};
int launch_gemm(const GemmCall& g, cudaStream_t s, int64_t* launches);
void gemm_force_config(int cfg);  // -1: heuristic (default); 0..4: force a DMMA tile configuration; 100+S: Ozaki, S slices

// ---- fp64 GEMM on the INT8 tensor pipe (ozaki.cu: tcgen05.mma.kind::i8 + TMEM + TMA, error-free splitting) --------
// on < 0 / slices == 0 / min_* <= 0 leave the respective setting unchanged.  Defaults come from the environment:
// GPC_OZAKI (0/1), GPC_OZAKI_SLICES (2..8), GPC_OZAKI_MIN_MN, GPC_OZAKI_MIN_K.
void oz_configure(int on, int slices, int64_t min_mn, int64_t min_k);
bool oz_wants(const GemmCall& g);
int oz_slices();                  // configured number of slices  // true when launch_gemm should route this call to launch_gemm_ozaki
int launch_gemm_ozaki(const GemmCall& g, cudaStream_t s, int64_t* launches, int slices = 0);  // 0: configured
void oz_release_device(int dev);   // frees the per-device slice workspace
void oz_set_debug(long long* dev_stamps, int cta);  // clock64 phase stamps of one CTA of the next GEMM launches (null: off)
// ---- block-cyclic one-sweep mode (dist.cu): panels pre-sliced into "slots", staircase update of the local matrix ----
struct OzCycMaps {   // the two tensor maps (A: 128-row boxes, B: 64-row boxes) over one slot buffer
  alignas(64) unsigned char opaque[2 * 128 + 128];
  const uint8_t* slots;
  int nb, S;
};
struct OzCycGrid { int P, Q, p, q; };
size_t oz_slot_bytes(int nb, int S);   // bytes of one slot: S planes of nb x nb int8 + nb row scales
int oz_slice_to_slots(const double* g, int64_t ld, bool kc, int64_t R, int nb, int S, int* emax_scratch, uint8_t* slots,
                      int slot0, int slot_stride, cudaStream_t s, int64_t* launches);
int oz_cyc_maps(OzCycMaps* out, const uint8_t* slots, int nslots, int nb, int S);
int launch_oz_cyc_update(const OzCycMaps& maps, const OzCycGrid& gr, int kstep, int skip_lo, int skip_hi, double* C,
                         int64_t ldc, int64_t r0, int64_t m, int64_t c0, int64_t n, int* errflag, cudaStream_t s,
                         int64_t* launches);

// potrf of one TILE x TILE diagonal block, in place (lower); also writes the inverse of the factor into
// Dinv (TILE x TILE, ld TILE, upper part zero), adds 2*sum(log diag) to *logdet and records the first
// non-positive pivot (1-based global order = base + j + 1) in *info if *info == 0.
int launch_potrf_leaf(double* A, int64_t lda, double* Dinv, int* info, int base, int64_t nvalid, double* logdet,
                      cudaStream_t s, int64_t* launches, double* Wd = nullptr, int64_t ldw = 0);
                      // Wd != null: the inverse of the factor is also written there with leading dimension ldw
void leaf_set_debug(long long* dev_stamps);  // phase time stamps of the next leaf launches (null: off)
// misc elementwise / reductions
int launch_copy_lower(const double* src, int64_t lds, double* dst, int64_t ldd, int64_t n, cudaStream_t s,
                      int64_t* launches);                       // dst lower(+diag) = src lower
int launch_mirror_lower(double* A, int64_t lda, int64_t n, cudaStream_t s, int64_t* launches);  // upper := lower'
int launch_copy_block(const double* src, int64_t lds, double* dst, int64_t ldd, int64_t m, int64_t n, double scale,
                      cudaStream_t s, int64_t* launches);
int launch_transpose(const double* src, int64_t lds, double* dst, int64_t ldd, int64_t m, int64_t n, cudaStream_t s,
                     int64_t* launches);                        // dst(n x m) = src(m x n)'
int launch_add_diag(double* A, int64_t lda, int64_t n, double v, cudaStream_t s, int64_t* launches);
int launch_zero_upper(double* A, int64_t lda, int64_t n, cudaStream_t s, int64_t* launches);
int launch_set_identity_pad(double* A, int64_t lda, int64_t n, int64_t np, cudaStream_t s, int64_t* launches);
// y(n x d) = A(n x n, full) * x(n x d); `part` is scratch of symm_chunks(n) * 4 * round_up(n, TILE) doubles
int symm_chunks(int64_t n);
int launch_symm_small(const double* A, int64_t lda, const double* x, int64_t ldx, double* y, int64_t ldy, int64_t n,
                      int d, double* part, cudaStream_t s, int64_t* launches);
// y = W x / W' x for a lower block-triangular W (block-upper part never read); part: scratch as for launch_symm_small
int launch_trmv_lower(const double* W, int64_t ldw, bool trans, const double* x, int64_t ldx, double* y, int64_t ldy,
                      int64_t n, int d, double* part, cudaStream_t s, int64_t* launches);
int launch_dot(const double* x, const double* y, int64_t n, double* out, cudaStream_t s, int64_t* launches);

// inverse of the lower-triangular TILE x TILE block at A (no factorisation): Dinv as above
int launch_trtri_leaf(const double* A, int64_t lda, double* Dinv, cudaStream_t s, int64_t* launches);

// ---- recursive blocked algorithms (api.cu) ---------------------------------------------------------------
struct GemmProf {  // optional per-launch timing of the DMMA GEMM kernel (bench.py roofline)
  std::vector<cudaEvent_t> ev;  // pairs
  std::vector<double> flops;    // executed flops of each launch
  struct Rec { int64_t m, n, k; int lower, ozaki; };
  std::vector<Rec> recs;        // shape and engine of each launch
  size_t used = 0;
};
struct Fork {  // side streams + events for fork/join concurrency inside the recursions (owned by the context)
  std::vector<cudaStream_t> side;
  std::vector<cudaEvent_t> ev;
  size_t next_s = 0, next_e = 0;
  // products rotate over all but the last four streams; those are kept for the short panel copies, which the main
  // stream waits for and which must never queue behind a long product
  cudaStream_t stream() { return side[next_s++ % (side.size() - 4)]; }
  cudaStream_t copy_stream() { return side[side.size() - 4 + next_c++ % 4]; }
  size_t next_c = 0;
  cudaEvent_t event() { return ev[next_e++ % ev.size()]; }
};
// Row-block pipeline of the TOP level of potrf_inv_rec when the inverse follows (gpc_eval): the second half of the
// matrix is factored as two diagonal nodes a', b'; as soon as a diagonal node is done, the rows of W = L^-1 of that row
// block and their contribution to K^-1 = W'W are queued on a low-priority bulk stream, where they run under the serial
// chain of the next diagonal node instead of after the whole factorisation (api.cu)
struct TopPipe {
  bool on = false, y_queued = false;
  int64_t n1 = 0, n2 = 0, h1 = 0;  // top-level split; rows of a' (split of the second half)
  double* Kinv = nullptr;
  int64_t ldo = 0;
  double* T = nullptr;             // top-level T = L21 W11 (n2 x n1, ld n2)
  cudaStream_t bulk = nullptr;
  cudaEvent_t e_w11 = nullptr, e_half = nullptr, e_bulk = nullptr, t_ready = nullptr;
};
// Look-ahead of the TOP level of potrf_inv_rec (gpc_eval): A11 is factored as two diagonal nodes a, b; as soon as a is
// done, the a-columns of L21 = A21 W11' and their share of the trailing update A22 -= L21 L21' run on the bulk stream
// under the serial chain of b, and only the b-columns are left on the critical path after A11 (api.cu)
struct TopFront {
  bool on = false, queued = false;
  int64_t n1 = 0, n2 = 0, h = 0, lda = 0;
  double *A21 = nullptr, *A22 = nullptr;
  const double* TL = nullptr;   // copy of A21 (n2 x n1, ld n2)
  cudaEvent_t e_copy = nullptr, e_a = nullptr, e_x2 = nullptr;
};
struct Dense {
  TopPipe* tp = nullptr;
  TopFront* tf = nullptr;
  int sm_limit = 0;    // copied into every product issued through this view (GemmCall::sm_limit)
  cudaStream_t bulk = nullptr;  // low-priority stream for large products that run next to the chain (null: side streams)
  GemmProf* prof = nullptr;
  Fork* fk = nullptr;  // null: everything on `s`
  cudaStream_t s;
  int64_t* launches;
  double* Dinv;    // n_total x TILE : inverse of diagonal block b at Dinv + b*TILE*TILE
  int* info;       // device
  double* logdet;  // device
  double* W;       // workspace for the inverse (>= (n/2 + TILE)^2 doubles)
  int64_t nvalid;  // rows below this index are real data (identity padding beyond)
  // potrf_inv_rec / inverse_from_W (factor + explicit triangular inverse, every solve is one large GEMM):
  double* Winv = nullptr;  // n_total x n_total, lower block triangle valid: W = L^-1
  int64_t ldw = 0;
  double* tmpL = nullptr;  // >= (n/2 + TILE) * (n/2) doubles: out-of-place result of a panel solve
  double* Tpool = nullptr; // >= tspace(n_total) doubles: T = L21 W11 per recursion depth
  double* TLpool = nullptr; // >= tspace(n_total) doubles: copy of A21 per recursion depth (taken on a side stream)
  cudaEvent_t* top_t_ready = nullptr;  // where the event of an early top-level T product is left for inverse_from_W
  // inverse of the diagonal block of the factor at row/column dbase: the diagonal block of W when W is kept, else Dinv
  const double* dinv_blk(int64_t dbase, int64_t* ld) const {
    if (Winv) {
      *ld = ldw;
      return Winv + dbase + dbase * ldw;
    }
    *ld = TILE;
    return Dinv + dbase * TILE;
  }
};
size_t potrf_inv_tspace(int64_t n);  // doubles of Tpool needed for an n x n factorisation
// in-place lower Cholesky of A (n x n block at row/column `base` of the matrix) AND W = L^-1 of the block into
// d.Winv; defer_top: the top-level W21 (3/4 of the inverse's flops) is left to inverse_from_W
int potrf_inv_rec(const Dense& d, double* A, int64_t lda, int64_t n, int64_t base, double* T, bool defer_top,
                  size_t tl_off = 0);
// completes W (if deferred) and forms Out = W' W = (L L')^-1, full symmetric
int inverse_from_W(const Dense& d, const double* L, int64_t ldl, int64_t n, double* Out, int64_t ldo, bool deferred_top);
int trsm_rlt(const Dense& d, double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl, int64_t n, int64_t dbase);
int trsm_rln(const Dense& d, double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl, int64_t n, int64_t dbase);
int potrf_rec(const Dense& d, double* A, int64_t lda, int64_t n, int64_t base, cudaEvent_t pending = nullptr);
int potri_rec(const Dense& d, const double* L, int64_t ldl, int64_t n, double* Out, int64_t ldo, int64_t dbase,
              double* W = nullptr);
size_t potri_workspace(int64_t n);  // doubles needed by potri_rec for an n x n factor

// ---- GP layer (gpkern.cu) -----------------------------------------------------------------------------
// local part of a block-cyclic N x N matrix (multi-GPU path): nb x nb blocks over a P x Q grid, this rank (p, q) holds
// local block (il, jl) = global block (il P + p, jl Q + q) in one ML x NL column-major matrix
struct CycMap {
  int on, nb, P, Q, p, q;
  int64_t ML, NL;
};
// cross-covariance mode of the gradient pass: sum_ij Cg[i,j] dk(X_i, X2_j)/dtheta and, into gX, d/dX_i
// (CKern::getGradParams(g, X, X2, covGrad) CKern.h:199-213 and overrides; getGradX CKern.h:68-74)
struct GradCross {
  int on;
  const double* X2;
  int64_t ldx2, n2;
};
int launch_kbuild_cyc(const KSpec& ks, const double* X, int64_t ldx, int64_t n, double* T, int64_t ldt, const CycMap& cm,
                      double jitter, cudaStream_t s, int64_t* launches);
int launch_symv_cyc(const double* T, int64_t ldt, const CycMap& cm, const double* x, int64_t ldx, int d, double* y,
                    int64_t ldy, cudaStream_t s, int64_t* launches);
// K (np x np, ld ldk): lower-triangle tiles of the kernel matrix of X (n valid rows, padded part = identity)
int launch_kbuild(const KSpec& ks, const double* X, int64_t ldx, int64_t n, int64_t np, double* K, int64_t ldk,
                  cudaStream_t s, int64_t* launches);
// Kc (n1p x n2p, ld ldk) = k(X1_i, X2_j) (computeElement semantics: white = 0), zero in the padding
int launch_kcross(const KSpec& ks, const double* X1, int64_t ldx1, int64_t n1, int64_t n1p, const double* X2,
                  int64_t ldx2, int64_t n2, int64_t n2p, double* Kc, int64_t ldk, cudaStream_t s, int64_t* launches,
                  int64_t col0 = -1);  // col0 >= 0: column block of the square training matrix (diag + identity pad)
int launch_kdiag(const KSpec& ks, const double* X, int64_t ldx, int64_t n, double* out, cudaStream_t s,
                 int64_t* launches);
// gradient pass.  mode 0: covGrad = -1/2 (dout*Kinv - alpha alpha')   (Kinv lower triangle read only)
//                 mode 1: covGrad = Cg (caller supplied, symmetric; lower triangle read only)
// partial (gridDim x nparams) receives per-CTA sums, reduced by launch_reduce_partials into g (nparams).
// gX (n x D, ld ldgx) accumulated with atomics when non-null (must be zeroed by the caller).
int launch_grad(const KSpec& ks, const double* X, int64_t ldx, int64_t n, int64_t np, const double* Cg, int64_t ldc,
                const double* alpha, int64_t lda, int dout, int mode, double* partial, int max_ctas, double* g,
                double* gX, int64_t ldgx, cudaStream_t s, int64_t* launches, int64_t col0 = -1, int64_t ncols = 0,
                const CycMap* cyc = nullptr, const GradCross* cross = nullptr);
                // col0 >= 0: only the lower-triangle tiles of columns [col0, col0+ncols) (multi-GPU column ownership)
                // cyc: Cg is the LOCAL part of a block-cyclic matrix (ld ldc); X, alpha are the full (replicated) inputs
// out[i] -= / = helpers for the posterior
int launch_row_sqnorm_sub(const double* V, int64_t ldv, int64_t rows, int64_t cols, const double* kdiag, double* var,
                          cudaStream_t s, int64_t* launches);  // var[i] = kdiag[i] - sum_j V[i,j]^2
int launch_gemv_rows(const double* A, int64_t lda, int64_t rows, int64_t cols, const double* x, int64_t ldx, int d,
                     double* y, int64_t ldy, cudaStream_t s, int64_t* launches);  // y(rows x d) = A(rows x cols) x

}  // namespace gpc
