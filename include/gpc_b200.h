/* gpc_b200.h -- C ABI of libgpc_b200.so: the B200 (sm_100a) implementation of GPc's exact-GP hot path.
 *
 * The reference (SheffieldML/GPc) has no C ABI of its own: its seams are (1) the Fortran BLAS/LAPACK
 * symbols declared in lapack.h:17-232 and (2) the C++ classes CMatrix / CKern / CGp / CGplvm.  This header
 * is what a maintainer binds behind those classes; every entry point cites the
 * reference code it replaces.  Conventions:
 *   - plain pointers and sizes only; all matrices column-major fp64 exactly like CMatrix (CMatrix.h:268);
 *   - pointers are HOST pointers unless the name ends in _dev;
 *   - return value: 0 ok; >0 LAPACK-style numerical info (order of the first non-positive pivot, as
 *     dpotrf_'s info, CMatrix.cpp:375-378); <0 usage / CUDA error, text via gpc_last_error();
 *   - no C++ exception crosses this boundary; the C++ shim turns info>0 into ndlexceptions::MatrixNonPosDef;
 *   - a gpc_ctx is single-threaded like a CGp object (mutable caches, CGp.h:365-447); several contexts on
 *     different threads / devices are fine.
 * There is NO CPU fallback: every call fails with GPC_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef GPC_B200_H
#define GPC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden: only this ABI is exported */
#endif

#define GPC_OK 0
#define GPC_ERR_ARG (-1)
#define GPC_ERR_CUDA (-2)
#define GPC_ERR_STATE (-3)
#define GPC_ERR_NOMEM (-4)

/* Kernel component types in scope.  Parameter order per component is the reference's:
 *   WHITE    [variance]                              CKern.cpp:604-739
 *   BIAS     [variance]                              CKern.cpp:889-1024
 *   RBF      [inverseWidth, variance]                CKern.cpp:1028-1256
 *   RBFARD   [inverseWidth, variance, scale_1..D]    CKern.cpp:3158-3417
 *   MATERN32 [lengthScale, variance]                 CKern.cpp:1708-1952
 *   MATERN52 [lengthScale, variance]                 CKern.cpp:1955-2216
 *   LIN      [variance]                              CKern.cpp:2220-2383
 *   POLY     [weightVariance, biasVariance, variance] (+ fixed degree)   CKern.cpp:2641-2905
 * A list of components is CCmpndKern's sum (CKern.cpp:128-328). */
enum gpc_kern_type {
  GPC_KERN_WHITE = 0,
  GPC_KERN_BIAS = 1,
  GPC_KERN_RBF = 2,
  GPC_KERN_RBFARD = 3,
  GPC_KERN_MATERN32 = 4,
  GPC_KERN_MATERN52 = 5,
  GPC_KERN_LIN = 6,
  GPC_KERN_POLY = 7
};
#define GPC_MAX_COMPONENTS 16
#define GPC_MAX_PARAMS 288

typedef struct gpc_kcomp {
  int type;             /* gpc_kern_type */
  int nparams;          /* must equal gpc_kern_nparams(type, D) */
  const double* params; /* NATURAL (untransformed) values, reference order */
  double degree;        /* POLY only (CPolyKern::degree, default 2) */
} gpc_kcomp;

/* transforms (CTransform.cpp:25-53 exp with +-36 clamp, :90-112 sigmoid with [eps,1-eps] clamp) */
enum gpc_transform { GPC_TRANS_NONE = 0, GPC_TRANS_EXP = 1, GPC_TRANS_SIGMOID = 2 };
int gpc_kern_nparams(int type, int D);
int gpc_kern_transform(int type, int param_index); /* which transform the reference attaches to this parameter */
double gpc_transform_atox(int transform, double a);
double gpc_transform_xtoa(int transform, double x);
double gpc_transform_gradfact(int transform, double x);

typedef struct gpc_ctx gpc_ctx;

const char* gpc_last_error(void);
int gpc_device_count(void);

/* ---- context: device-resident state of one CGp / CGplvm object --------------------------------------
 * Owns X (N x D), m (N x d), K, L (= LcholK, lower), K^-1 (= invK), alpha; replaces the four N^2 host
 * matrices CGp::initStoreage allocates (CGp.cpp:171-175, 178-246). */
int gpc_ctx_create(gpc_ctx** out, int device, int64_t Nmax, int Dmax, int dout_max);
int gpc_ctx_destroy(gpc_ctx* ctx);
/* run on a caller-owned CUDA stream (cudaStream_t as void*); NULL restores the context's own stream */
int gpc_ctx_set_stream(gpc_ctx* ctx, void* cuda_stream);
void* gpc_ctx_get_stream(gpc_ctx* ctx);
int gpc_ctx_sync(gpc_ctx* ctx);
/* number of kernel launches issued through this context since creation (bench.py's gpu_launches) */
int64_t gpc_ctx_launch_count(gpc_ctx* ctx);

/* X := host N x D (pX of CGp, CGp.h:353); M := host N x d, M = (y - bias)/scale i.e. CGp::updateM
 * (CGp.cpp:248-260) is done by the caller, or use gpc_set_Y to do it on the device. */
int gpc_set_X(gpc_ctx* ctx, const double* X, int64_t N, int D, int64_t ldx);
int gpc_set_M(gpc_ctx* ctx, const double* M, int64_t N, int d, int64_t ldm);
int gpc_set_Y(gpc_ctx* ctx, const double* Y, int64_t N, int d, int64_t ldy, const double* bias, const double* scale);

/* K := kernel matrix of X.  Replaces CGp::_updateK FTC (CGp.cpp:693-712) / CKern::compute(K,X) (CKern.h:128-144):
 * off-diagonal computeElement, diagonal diagComputeElement (white noise enters here only). */
int gpc_kern_build(gpc_ctx* ctx, const gpc_kcomp* comps, int ncomp);
/* K(X, Xs) -> host Ks (N x Ns, ld ldk).  Replaces CKern::compute(K,X,X2) (CKern.h:146-157); white contributes 0. */
int gpc_kern_cross(gpc_ctx* ctx, const gpc_kcomp* comps, int ncomp, const double* Xs, int64_t Ns, int64_t ldxs,
                   double* Ks, int64_t ldk);
/* diag k(Xs_i, Xs_i) -> host (CKern::diagCompute, CKern.h:50-56) */
int gpc_kern_diag(gpc_ctx* ctx, const gpc_kcomp* comps, int ncomp, const double* Xs, int64_t Ns, int64_t ldxs,
                  double* kdiag);
/* K += jitter * I   (CMatrix::addDiag as used by jitChol, CMatrix.cpp:787) */
int gpc_add_diag(gpc_ctx* ctx, double jitter);
/* L := chol(K) lower, logdet = 2 sum log L_ii.  Replaces LcholK.chol()/potrf (CMatrix.cpp:371-403, dpotrf_
 * lapack.h:59-65) + logDet (CMatrix.cpp:404-412) + the zero-lower / Alg.513 transpose (CMatrix.cpp:391-395,
 * CGp.cpp:890).  *info: 0 or the 1-based order of the first non-positive pivot (returned value is the same). */
int gpc_potrf(gpc_ctx* ctx, int* info, double* logdet);
/* jitChol (CMatrix.cpp:767-804) on the device-resident K: same jitter schedule, K mutated the same way.
 * *jitter receives the value the reference returns. Returns >0 (info) if it gives up like the reference does. */
int gpc_jitchol(gpc_ctx* ctx, int max_tries, double* jitter, double* logdet);
/* alpha := L^-T L^-1 m (CGp::updateAlpha, CGp.cpp:469-484); *quad = sum_j m_j' alpha_j */
int gpc_solve_alpha(gpc_ctx* ctx, double* quad);
/* K^-1 from L (CMatrix::pdinv, CMatrix.cpp:421-432, dpotri_ lapack.h:67-73), full symmetric on the device */
int gpc_inverse(gpc_ctx* ctx);
/* alpha := K^-1 m via the explicit inverse (dsymv_, as CGp::logLikelihood does, CGp.cpp:924-930) */
int gpc_alpha_from_inverse(gpc_ctx* ctx, double* quad);
/* gparams[P] = sum_j sum_ik covGrad_j[i,k] dK[i,k]/dtheta, covGrad_j = -1/2 (K^-1 - alpha_j alpha_j')
 * (CGp::updateCovGradient CGp.cpp:666-679 + CKern::getGradParams, e.g. CKern.cpp:1204-1241) WITHOUT ever
 * materialising covGrad.  NATURAL-parameter gradients, component order; multiply by gpc_transform_gradfact for
 * getGradTransParams (CKern.cpp:50-63).  gX (N x D, ld N) optional: d ll / d X as CGplvm accumulates it
 * (CGplvm.cpp:569-603), without the latent prior term.  Needs gpc_inverse + an alpha. */
int gpc_grad(gpc_ctx* ctx, const gpc_kcomp* comps, int ncomp, double* gparams, double* gX);
/* generic sum_ik covGrad[i,k] dK[i,k]/dtheta for a caller-supplied HOST covGrad (N x N, symmetric):
 * CKern::getGradParams(g, X, covGrad) (CKern.h:187-197 and overrides) -- used by the kernel unit tests */
int gpc_kern_grad(gpc_ctx* ctx, const gpc_kcomp* comps, int ncomp, const double* covGrad, int64_t ldc,
                  double* gparams, double* gX);
/* the cross-covariance form: gparams[P] = sum_ij covGrad2[i,j] dk(X_i, X2_j)/dtheta for HOST X2 (N2 x D) and covGrad2
 * (N x N2) -- CKern::getGradParams(g, X, X2, covGrad) (CKern.h:199-213 and overrides, exercised by testKern.cpp:306-325,
 * "g4") -- and, if gX != NULL (N x D, ld N), gX[i,:] = sum_j covGrad2[i,j] dk(X_i, X2_j)/dX_i, the covGrad-weighted
 * CKern::getGradX (CKern.h:68-74; testKern.cpp:327-362, "G2").  computeElement semantics: white contributes nothing. */
int gpc_kern_grad_cross(gpc_ctx* ctx, const gpc_kcomp* comps, int ncomp, const double* X2, int64_t N2, int64_t ldx2,
                        const double* covGrad2, int64_t ldc, double* gparams, double* gX);
/* posterior at Xs (Ns x D): mu (Ns x d, ld Ns), var (Ns x d, ld Ns) in the space of m (caller applies
 * scale/bias, CGp.cpp:561-573, 622).  Replaces CGp::posteriorMeanVar (CGp.cpp:642-663): K(X,Xs) build,
 * mu = K*' alpha, var = k(x*,x*) - |L^-1 k*|^2 (dtrsm_, CGp.cpp:603-606).  var may be NULL. */
int gpc_posterior(gpc_ctx* ctx, const gpc_kcomp* comps, int ncomp, const double* Xs, int64_t Ns, int64_t ldxs,
                  double* mu, double* var);

/* One full evaluation = what one SCG step asks of CGp (COptimisable.cpp:309-349):
 * K build -> jitChol -> K^-1 -> alpha -> ll terms -> gradient.  out[0]=logdet, out[1]=quad, out[2]=jitter used.
 * flags: bit0 = also compute gX (needs gX != NULL). Single host sync at the end.
 * Afterwards the context holds L, W = L^-1, alpha and the LOWER triangle of K^-1 (gpc_download(GPC_MAT_KINV) returns the
 * full symmetric matrix); K was factored in the buffer it was built in and is rebuilt from this evaluation's kernel by
 * whoever needs it next (gpc_download(GPC_MAT_K), gpc_potrf, gpc_jitchol, gpc_add_diag).  alpha = W'(W m). */
int gpc_eval(gpc_ctx* ctx, const gpc_kcomp* comps, int ncomp, int flags, double* out, double* gparams, double* gX);

enum gpc_which { GPC_MAT_K = 0, GPC_MAT_L = 1, GPC_MAT_KINV = 2, GPC_MAT_ALPHA = 3, GPC_MAT_M = 4 };
/* copy a device-resident matrix to the host (N x N or N x d).  K and K^-1 come back full symmetric, L lower
 * with a zero strict upper triangle (LcholK after .trans(), CGp.cpp:890). */
int gpc_download(gpc_ctx* ctx, int which, double* dst, int64_t ld);
/* phase timings of the last gpc_eval in milliseconds (CUDA events on the context's stream):
 * [0] K build [1] potrf [2] inverse [3] alpha+reductions [4] gradient [5] total */
int gpc_last_timings(gpc_ctx* ctx, double* ms6);
/* host milliseconds the last gpc_eval spent queueing work before its single synchronisation (launch-bound check) */
int gpc_last_enqueue_ms(gpc_ctx* ctx, double* ms);

/* profiling mode (bench.py roofline): bracket every DMMA GEMM launch of gpc_eval with CUDA events on the context's
 * stream; after an evaluation gpc_last_gemm_profile reports the summed kernel time, launch count and executed flops */
int gpc_ctx_set_profile(gpc_ctx* ctx, int on);
int gpc_last_gemm_profile(gpc_ctx* ctx, double* total_ms, int64_t* count, double* flops);
/* the same split by engine: out8 = DMMA [ms, flops, launches, 0], Ozaki [ms, fp64-equivalent flops, launches, int8 ops
 * executed].  With GPC_PROF_DUMP=<file> in the environment every profiled GEMM is also written there (shape, engine, ms). */
int gpc_last_gemm_profile_split(gpc_ctx* ctx, double* out8);

/* ---- CMatrix level: drop-ins for the lapack.h calls CMatrix makes (host pointers, LAPACK argument meaning,
 *      scalars by value, 64-bit dimensions).  Each stages through device memory. ---------------------------*/
/* dpotrf_ (lapack.h:59-65; CMatrix::potrf CMatrix.cpp:371-379) */
int gpc_dpotrf(int device, char uplo, int64_t n, double* A, int64_t lda, int* info);
/* dpotri_ (lapack.h:67-73; CMatrix::potri CMatrix.cpp:414-420): only the `uplo` triangle is written */
int gpc_dpotri(int device, char uplo, int64_t n, double* A, int64_t lda, int* info);
/* dtrsm_ (lapack.h:214-222; CMatrix::trsm CMatrix.cpp:272-295) */
int gpc_dtrsm(int device, char side, char uplo, char transa, char diag, int64_t m, int64_t n, double alpha,
              const double* A, int64_t lda, double* B, int64_t ldb);
/* dsyrk_ (lapack.h:196-203; CMatrix::syrk CMatrix.cpp:297-322) */
int gpc_dsyrk(int device, char uplo, char trans, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
              double beta, double* C, int64_t ldc);
/* dgemm_ (lapack.h:186-194; CMatrix::gemm CMatrix.cpp:205-247) */
int gpc_dgemm(int device, char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
              int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc);
/* dsyr_ (lapack.h:154-160; CMatrix::syr CMatrix.h:526-533): A := alpha x x' + A on the `uplo` triangle */
int gpc_dsyr(int device, char uplo, int64_t n, double alpha, const double* x, int64_t incx, double* A, int64_t lda);
/* dsymv_ (lapack.h:152-160; CMatrix::symv CMatrix.cpp:127-203) */
int gpc_dsymv(int device, char uplo, int64_t n, double alpha, const double* A, int64_t lda, const double* x,
              double beta, double* y);

/* ---- sparse approximations of CGp: DTC, DTCVAR, FITC with M inducing inputs --------------------------
 * CGp.cpp:713-735 (K_uu / K_uf / diag K build), :766-861 (updateAD), :939-988 (logLikelihood), :1146-1223, :1244-1413
 * (gradients), :490-521, :584-599 (posterior).  approx takes the reference's enum values (CGp.h:13-19): 1 DTC, 2 FITC,
 * 4 DTCVAR.  One evaluation = K_uu, K_uf (and k_ii) build, two M x M factorisations with explicit inverses, the
 * O(N M^2) products on the GEMM engines, and the gradient passes over K_uu (symmetric) and K_uf (cross); nothing N x N. */
typedef struct gpc_sparse gpc_sparse;
int gpc_sparse_create(gpc_sparse** out, int device, int approx, int64_t N, int M, int D, int dout);
int gpc_sparse_destroy(gpc_sparse* h);
/* X (N x D) and m = (y - bias) / scale (N x dout, CGp::updateM CGp.cpp:248-260), HOST pointers */
int gpc_sparse_set_data(gpc_sparse* h, const double* X, int64_t ldx, const double* M, int64_t ldm);
/* Xu: M x D inducing inputs (host), beta: the noise precision.  out[0] = log-likelihood exactly as CGp::logLikelihood
 * returns it (with the reference's constants: -d N/2 log 2pi, counted twice for FITC, CGp.cpp:963 + :1012; without priors and
 * learnt-scale terms), out[1] = log|Sigma|, out[2] = sum_j m_j' Sigma^-1 m_j, out[3] = sum_i (k_ii - q_ii),
 * out[4], out[5] = jitter the jitChol schedule added to K_uu and to A (0 normally; CGp.cpp:777, 830 call jitChol on A).
 * gparams[P]: NATURAL kernel-parameter gradients, component order (multiply by gpc_transform_gradfact); gXu (M x D, ld M):
 * d ll / d X_u; *gbeta: d ll / d beta (the optimiser's log-beta gradient is beta times it, CGp.cpp:1073-1076).
 * Returns 0, >0 = info of a factorisation that stayed non positive definite through the jitChol schedule, <0 error. */
int gpc_sparse_eval(gpc_sparse* h, const gpc_kcomp* comps, int ncomp, const double* Xu, int64_t ldxu, double beta,
                    double* out, double* gparams, double* gXu, double* gbeta);
/* mu (Ns x dout, ld Ns) and var (Ns; the same for every output) at HOST Xs (Ns x D), in the space of m; needs a
 * gpc_sparse_eval at the current parameters.  var may be NULL. */
int gpc_sparse_posterior(gpc_sparse* h, const gpc_kcomp* comps, int ncomp, const double* Xs, int64_t Ns, int64_t ldxs,
                         double* mu, double* var);
int64_t gpc_sparse_launch_count(gpc_sparse* h);

/* ---- multi-GPU: K, its factor and K^-1 sharded over a P x Q process grid ------------------------------
 * For N beyond one GPU's memory (the reference cannot even index N = 65536: unsigned int nrows*ncols, CMatrix.cpp:654,
 * CMatrix.h:232).  The N x N matrix is cut into nb x nb blocks dealt 2-D block-cyclically (rank = p*Q + q owns block (i, j)
 * iff i % P == p and j % Q == q); each rank stores only its blocks on / below the diagonal (N^2 / (P Q) doubles) and
 * K -> K^-1 happens in place in ONE right-looking sweep that fuses dpotrf_ and dpotri_ (CMatrix.cpp:371-432,
 * lapack.h:59-73): per step one panel (diagonal-block factor + its inverse, panel products) is broadcast, already split
 * into the int8 planes the tensor-core engine consumes, and every rank applies the same rank-nb update to all its blocks.
 * log det, alpha = K^-1 m and the gradient partial sums are all-reduced.  Same quantities as gpc_eval.
 * Two back-ends:
 *   gpc_dist_create_nccl : one PROCESS per GPU (mpirun-style launchers).  Rank 0 calls gpc_dist_unique_id and the
 *                          launcher hands the 128 bytes to every rank (any transport); collectives = ncclBroadcast /
 *                          ncclAllReduce on the library's own communicator (libnccl.so.2 is loaded at run time).
 *   gpc_dist_create_local: ONE process driving ndev devices, a worker thread per device, peer copies over NVLink
 *                          ordered by CUDA events -- SURVEY 8(b)'s gpc_ctx_create(devices, ndev, ...).  A device may be
 *                          listed more than once (the block-cyclic logic can be exercised on a single GPU).
 * Every call on a gpc_dist is collective in the NCCL mode (all ranks call it with the same arguments). */
typedef struct gpc_dist gpc_dist;
int gpc_dist_unique_id(void* id128);
int gpc_dist_create_nccl(gpc_dist** out, int device, int rank, int world, const void* id128, int P, int Q, int64_t N, int D,
                         int dout, int nb);
int gpc_dist_create_local(gpc_dist** out, const int* devices, int ndev, int P, int Q, int64_t N, int D, int dout, int nb);
int gpc_dist_destroy(gpc_dist* h);
/* X (N x D) and m (N x dout), HOST pointers, replicated to every rank (X is at most a few MB: SURVEY 8(e)) */
int gpc_dist_set_data(gpc_dist* h, const double* X, int64_t ldx, const double* M, int64_t ldm);
/* one evaluation: out[0] = logdet, out[1] = quad, out[2] = jitter used (jitChol schedule, CMatrix.cpp:767-804: K is
 * rebuilt with the accumulated jitter and the sweep repeated); gparams as gpc_eval.  Returns 0, >0 = info, <0 error. */
int gpc_dist_eval(gpc_dist* h, const gpc_kcomp* comps, int ncomp, double* out, double* gparams);
/* the blocks of K^-1 this process holds, scattered into a host N x N matrix (both triangles); other entries untouched.
 * Test hook: with the local back-end the whole matrix comes back, with NCCL each process fills in its own blocks. */
int gpc_dist_download_kinv(gpc_dist* h, double* dst, int64_t ld);
/* out8: [ranks, steps (= block rows), bytes of this rank's local matrix, bytes of its two panel buffers, bytes broadcast
 * per step, kernel launches so far, nb, back-end (0 local, 1 nccl)]; ms5: phases of the last evaluation on this rank
 * [K build, sweep, alpha, gradient, reductions] */
int gpc_dist_info(gpc_dist* h, int64_t* out8, double* ms5);
/* host-only (no GPU needed): what `rank` of a P x Q grid does at step k of the sweep for an N x N problem in nb blocks.
 * out12 = [block rows NBt, local block rows, local block columns, owns part of block column k, owns part of block row k,
 * owner of block (k,k), first local block row of its column part (-1), its first slot, number of local block columns of
 * its row part (-1), first local block row / number of block columns of the look-ahead strips, block row+column the bulk
 * update leaves out (-1: none)]; producers (NBt ints, may be NULL) = the rank that produces each slot of panel k */
int gpc_dist_plan(int P, int Q, int rank, int64_t N, int nb, int k, int* out12, int* producers);

/* ---- measurement helpers (bench.py) ----------------------------------------------------------------- */
/* register-resident DMMA loop: measured fp64 tensor-pipe peak of this device in TFLOP/s */
int gpc_bench_dmma_peak(int device, double* tflops);
/* the same for the INT8 tensor pipe the large fp64 products run on (ozaki.cu): one CTA per SM issuing tcgen05.mma.kind::i8
 * M = 128, N = 64 nwide (nwide 1..4), K = 32 on shared-memory-resident operands into TMEM; *tops = 2 x MACs / s / 1e12.
 * The fp64-equivalent peak of the engine is tops / (S (S + 1) / 2) for S slices (36 int8 MMAs per fp64 MMA at S = 8). */
int gpc_bench_imma_peak(int device, int nwide, double* tops);
/* the same loop kept running for `seconds`: the SUSTAINED rate over the second half (under the clock the power cap leaves) --
 * the denominator for a kernel timed inside a long step, where the burst figure above is not reachable */
int gpc_bench_imma_peak_sustained(int device, int nwide, double seconds, double* tops);
/* C(n x n) -= A(n x k) A' (lower) on device scratch: the SYRK trailing update in isolation.  *ms per launch */
int gpc_bench_syrk(int device, int64_t n, int64_t k, int reps, double* ms);
/* the 128 x 128 diagonal-block kernel (Cholesky + inverse of the factor, N/128 times on the critical path) in
 * isolation: *us per launch; stamps (32 entries) = clock64 phase stamps of one launch relative to its start */
int gpc_bench_leaf(int device, int reps, double* us, long long* stamps);
/* clock64 phase stamps (8 entries, relative to the CTA's start) of CTA `cta` of one tensor-core GEMM m x n x k (kernel tuning):
 * start, barriers + TMEM ready, first operands landed, last MMA issued, C segment requested, accumulators complete,
 * epilogue stores issued, end */
int gpc_bench_oz_stamps(int device, int64_t m, int64_t n, int64_t k, int cta, long long* stamps8);
/* one GEMM shape on device scratch with a forced tile configuration (cfg 0..3, -1 = heuristic): kernel tuning */
int gpc_bench_gemm(int device, int64_t m, int64_t n, int64_t k, int a_kc, int b_kc, int lower, int cfg, int reps,
                   double* ms);

/* ---- the callers either side of the path ----------------------------------------------- */
/* sizes of the problem currently loaded into the context */
int gpc_ctx_dims(gpc_ctx* ctx, int64_t* N, int* D, int* dout);
/* Scaled conjugate gradients on the kernel hyper-parameters of an FTC GP: CGp::optimise -> COptimisable::scgOptimise
 * (COptimisable.cpp:246-396; same steps, same quirks: CGp.cpp:1537-1553).  comps[].params (NATURAL values, must point to
 * writable memory) are the start point and receive the result; the search runs in the reference's transformed space
 * (exp / sigmoid, CTransform.cpp:25-112).  trace (may be NULL, max_iters doubles) receives the objective
 * -log p(y|X,theta) after every iteration.  One device evaluation per distinct point: the reference's second
 * factorisation at an accepted point (COptimisable.cpp:333 then :347) is served from the first.
 * Returns 0, or gpc_eval's code if an evaluation failed (the parameters then hold the last accepted point). */
int gpc_gp_optimise_scg(gpc_ctx* ctx, gpc_kcomp* comps, int ncomp, int max_iters, double param_tol, double obj_tol,
                        double* trace, int* iters_out, int* evals_out);
/* The same loop for any objective (the COptimisable interface, COptimisable.h:15-239): fn returns 0 and fills *obj and
 * grad[n] at w[n]; it is called once per distinct point.  w holds the start point and receives the result. */
typedef int (*gpc_objective_fn)(void* user, const double* w, int n, double* obj, double* grad);
int gpc_scg_minimise(gpc_objective_fn fn, void* user, double* w, int n, int max_iters, double param_tol, double obj_tol,
                     double* trace, int* iters_out, int* evals_out);
/* SVM-light files as CClctrl::readSvmlDataFile reads them (CClctrl.cpp:55-171): label first, then 1-based index:value
 * pairs separated by single spaces, '#' lines skipped, '\r' dropped; D = the largest index in the file.
 * gpc_svml_dims sizes the buffers; gpc_svml_read fills X (nrows x ncols, column-major, ld ldx, zero where absent) and y. */
int gpc_svml_dims(const char* path, int64_t* nrows, int* ncols);
int gpc_svml_read(const char* path, double* X, int64_t ldx, double* y, int64_t nrows, int ncols);

/* GP model files: the text format `gp learn` writes and `gp display / gnuplot / relearn` read (CGp::writeParamsToStream /
 * readParamsFromStream CGp.cpp:1605-1682; CStreamInterface CNdlInterfaces.h:21-175; CMatrix CMatrix.cpp:1057-1097,
 * 1158-1172; CKern CKern.cpp:15-26, 94-126, 2668-2705, 4192-4278; CNoise CNoise.cpp:275-305, 1813-1836).  Exact FTC models
 * with the device kernels; X and y are not part of the file (gp.cpp:486-490 reloads them).  Files written here are
 * byte-identical to the reference's for the same model (after the "# comment" line) and files are read the way the
 * reference reads them, quirks included: nested versions and matrix entries are hexfloat text, integer values are written
 * as integers, and an entry WITHOUT a '.' is read with atoi (CMatrix.cpp:1081-1085) -- 0.25, written "0x1p-2", comes back
 * as 0.  gpc_gp_model_check_roundtrip counts the values of a model that would be lost that way (*first_lost = the first).
 * A file with priors is rejected, as the reference's own reader rejects it (CDist.cpp:4-10 vs 338-357). */
#define GPC_MODEL_MAX_OUT 256
typedef struct gpc_kern_spec {       /* a kernel object as the stream holds it (readKernFromStream, CKern.cpp:4192-4259) */
  int top_is_cmpnd;                  /* 1: CCmpndKern of ncomp components; 0: a single kernel object (ncomp = 1) */
  int input_dim;
  int ncomp;
  int type[GPC_MAX_COMPONENTS];      /* gpc_kern_type */
  int nparams[GPC_MAX_COMPONENTS];
  double degree[GPC_MAX_COMPONENTS]; /* POLY only */
  double params[GPC_MAX_PARAMS];     /* NATURAL values, component order (CKern::getParams) */
} gpc_kern_spec;
typedef struct gpc_noise_spec {      /* a noise object as the stream holds it (CNoise.cpp:275-305, 1813-1836) */
  char type[16];                     /* "gaussian" for gp learn, "scale" for gplvm learn */
  int output_dim, nparams;
  double params[2 * GPC_MODEL_MAX_OUT + 8]; /* gaussian: bias_1..bias_d, sigma2; scale: bias_1..bias_d, scale_1..scale_d */
} gpc_noise_spec;
typedef struct gpc_gp_model {
  int64_t num_data;
  int input_dim, output_dim;
  int approx_type;         /* CGp::FTC = 0 (field "sparseApproximation"); anything else is rejected */
  unsigned int num_active; /* written as is (the CLI leaves 0 or 4294967295 for FTC) */
  int learn_scale, learn_bias;
  gpc_kern_spec kern;
  double scale[GPC_MODEL_MAX_OUT], bias[GPC_MODEL_MAX_OUT];
  gpc_noise_spec noise;
} gpc_gp_model;
int gpc_gp_model_read(const char* path, gpc_gp_model* out);
int gpc_gp_model_write(const char* path, const gpc_gp_model* model, const char* comment);
int gpc_gp_model_check_roundtrip(const gpc_gp_model* model, int* nlost, double* first_lost);

/* GP-LVM model files (CGplvm::writeParamsToStream / readParamsFromStream CGplvm.cpp:761-898, writeGplvmToFile :908-921):
 * header fields, kernel, CScaleNoise, then one text row per data point holding its d targets, q latent coordinates and
 * (optionally) an integer label -- all hexfloat, all read back with atof (no atoi rule here).  Plain GP-LVMs only: files
 * with back constraints or a dynamics kernel are rejected.  Y (num_data x output_dim) and X (num_data x latent_dim) are
 * caller-owned column-major buffers; labels (num_data ints) may be NULL.  gpc_gplvm_model_read with Y == X == NULL fills
 * the header only (sizes for the buffers). */
typedef struct gpc_gplvm_model {
  int64_t num_data;
  int output_dim, latent_dim;
  int latent_regularised, back_constrained, dynamics_learnt;
  int has_labels;
  gpc_kern_spec kern;
  gpc_noise_spec noise;
} gpc_gplvm_model;
int gpc_gplvm_model_read(const char* path, gpc_gplvm_model* out, double* Y, int64_t ldy, double* X, int64_t ldx, int* labels);
int gpc_gplvm_model_write(const char* path, const gpc_gplvm_model* model, const double* Y, int64_t ldy, const double* X,
                          int64_t ldx, const int* labels, const char* comment);

/* ---- fp64 GEMM engine selection -------------------------------------------------------------------- */
/* The dsyrk_/dgemm_ work below dpotrf_/dpotri_ (lapack.h:59-73) runs on one of two engines:
 *   DMMA  : mma.sync.m8n8k4.f64 (the fp64 tensor pipe, 37 TFLOP/s peak);
 *   OZAKI : error-free splitting into `slices` int8 planes multiplied exactly on tcgen05.mma.kind::i8 with int32
 *           accumulators in TMEM and recombined in fp64 (6 + 8(slices-1) bits below each row's largest entry; the
 *           default 8 slices = 62 bits, fp64 has 53).  Used for calls with m, n >= min_mn and k >= min_k that the
 *           engine cost model expects to be faster.
 * ozaki: 0 = DMMA only, 1 = Ozaki for large calls, -1 = leave; slices 2..8 (0 = leave); min_* <= 0 = leave.
 * Process-wide; initial values come from GPC_OZAKI, GPC_OZAKI_SLICES, GPC_OZAKI_MIN_MN, GPC_OZAKI_MIN_K. */
int gpc_set_gemm_engine(int ozaki, int slices, int64_t min_mn, int64_t min_k);
/* the number of slices currently configured */
int gpc_gemm_engine_slices(void);
/* One GEMM on HOST arrays in the padded device layouts (m, n multiples of 128, k of 128): A is m x k (ld m) or, with
 * a_kc, k x m (ld k); B likewise n x k / k x n; C m x n (ld m), updated in place: C = alpha op(A) op(B)' + beta C.
 * lower: bit 0 = only the tiles touching the lower triangle (m == n); bits 1-2 = triangular op(A) (1: zero for kk < i,
 * 2: zero for kk > i); bits 3-4 = triangular op(B) (1: zero for kk < j, 2: zero for kk > j) -- the zero part of a
 * triangular operand is skipped tile-wise and never read.
 * cfg: -1 heuristic, 0..4 DMMA tile configuration, 100+S Ozaki with S slices.  Engine parity tests. */
int gpc_gemm_check(int device, int64_t m, int64_t n, int64_t k, int a_kc, int b_kc, int lower, int cfg, double alpha,
                   double beta, const double* A, const double* B, double* C);

/* host-only test hook: launch boundaries of a product on the tensor-core engine that may hold at most `limit` SMs at a
 * time (tiles_m x tiles_n tiles of 128 x 64, lower: only tiles that intersect the lower triangle are live).  cuts == NULL:
 * returns the number of boundaries; else writes up to cap of them (first 0, last tiles_m * tiles_n). */
int gpc_oz_wave_cuts(int tiles_m, int tiles_n, int lower, int limit, long long* cuts, int cap);
/* The Ozaki splitting of one HOST operand (R rows = the m or n index, K deep; kc: k contiguous, ld K, else ld R):
 * slices_out receives S planes of R x K int8 (k contiguous), scale_out the R row scales 2^e_r, such that
 * x[r,k] = scale[r] * sum_p slices[p][r][k] * 2^-(8p+6) (p = 0..S-1) to 6+8(S-1) bits.  Test hook of the slicing kernel. */
int gpc_oz_slice_check(int device, int64_t R, int64_t K, int kc, int S, const double* X, signed char* slices_out,
                       double* scale_out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* GPC_B200_H */
