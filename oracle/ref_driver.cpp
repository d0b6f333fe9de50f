// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libgpc_b200.so).
//
// A flat extern "C" facade over the UNMODIFIED reference classes (CKern / CGp / CGplvm /
// CMatrix from /root/reference, compiled where they lie by oracle/build_ref.sh) so that
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm can
// drive the reference's own CPU path through ctypes.  All matrices are column-major fp64,
// exactly as CMatrix stores them (CMatrix.h:268).
//
// Kernel specification used by every entry point (mirrors include/gpc_b200.h):
//   ncomp component type codes (GPC_KERN_*), then the TRANSFORMED parameter vector of the
//   compound kernel in component order (what CKern::setTransParams takes, CTransform.h:281).
#include <cstring>
#include <time.h>
#include <string>
#include <vector>
#include "CKern.h"
#include "CGp.h"
#include "CGplvm.h"
#include "CNoise.h"
#include "CMatrix.h"
#include "CClctrl.h"
#include "COptimisable.h"

extern "C" {
void scipy_openblas_set_num_threads(int);
int scipy_openblas_get_num_threads(void);
}

namespace {
enum { K_WHITE = 0, K_BIAS = 1, K_RBF = 2, K_RBFARD = 3, K_MATERN32 = 4, K_MATERN52 = 5, K_LIN = 6, K_POLY = 7 };

static std::string g_err;  // (gnu++98: the reference does not compile as C++11)

CKern* makeComponent(int type, unsigned int D) {
  switch (type) {
    case K_WHITE: return new CWhiteKern(D);
    case K_BIAS: return new CBiasKern(D);
    case K_RBF: return new CRbfKern(D);
    case K_RBFARD: return new CRbfardKern(D);
    case K_MATERN32: return new CMatern32Kern(D);
    case K_MATERN52: return new CMatern52Kern(D);
    case K_LIN: return new CLinKern(D);
    case K_POLY: return new CPolyKern(D);
  }
  return 0;
}

// Build cmpnd(components...) exactly as gp.cpp:240-349 assembles it: addKern clones.
CCmpndKern* makeKern(int ncomp, const int* types, const double* tparams, unsigned int D) {
  CCmpndKern* kern = new CCmpndKern(D);
  for (int c = 0; c < ncomp; c++) {
    CKern* comp = makeComponent(types[c], D);
    if (!comp) { delete kern; return 0; }
    kern->addKern(comp);
    // NOTE: comp is deliberately leaked.  CRbfardKern's copy constructor assigns `scales = kern.scales`
    // (CKern.cpp:3176-3183) and CMatrix has no deep operator=, so clone and original share one buffer;
    // deleting either frees the other's storage.  gp.cpp keeps its originals alive too.
  }
  CMatrix tp(1, kern->getNumParams(), const_cast<double*>(tparams));
  kern->setTransParams(tp);
  return kern;
}
double now() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
}  // namespace

#define REF_TRY try {
#define REF_CATCH                                     \
  }                                                   \
  catch (ndlexceptions::Error & e) {                  \
    g_err = e.getMessage();                           \
    return -1;                                        \
  }                                                   \
  catch (std::exception & e) {                        \
    g_err = e.what();                                 \
    return -1;                                        \
  }                                                   \
  catch (...) {                                       \
    g_err = "unknown exception";                      \
    return -1;                                        \
  }

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }
void ref_set_threads(int n) { scipy_openblas_set_num_threads(n); }
int ref_get_threads() { return scipy_openblas_get_num_threads(); }

int ref_kern_nparams(int ncomp, const int* types, int D) {
  REF_TRY
  CCmpndKern kern((unsigned int)D);
  for (int c = 0; c < ncomp; c++) {
    CKern* comp = makeComponent(types[c], D);
    if (!comp) return -1;
    kern.addKern(comp);
  }
  return (int)kern.getNumParams();
  REF_CATCH
}

// untransformed parameter values after setTransParams (pins CTransform.cpp:25-53, 90-112)
int ref_kern_params(int ncomp, const int* types, const double* tparams, int D, double* params) {
  REF_TRY
  CCmpndKern* kern = makeKern(ncomp, types, tparams, D);
  if (!kern) return -1;
  for (unsigned int i = 0; i < kern->getNumParams(); i++) params[i] = kern->getParam(i);
  // kern (and its clones) leaked on purpose, see makeKern
  return 0;
  REF_CATCH
}

// K = kern.compute(K, X)  (CKern.h:128-144)
int ref_kern_compute(int ncomp, const int* types, const double* tparams, const double* X, int N, int D, double* K) {
  REF_TRY
  CCmpndKern* kern = makeKern(ncomp, types, tparams, D);
  if (!kern) return -1;
  CMatrix Xm(N, D, const_cast<double*>(X));
  CMatrix Km(N, N);
  kern->compute(Km, Xm);
  memcpy(K, Km.getVals(), sizeof(double) * (size_t)N * N);
  // kern (and its clones) leaked on purpose, see makeKern
  return 0;
  REF_CATCH
}

// K = kern.compute(K, X, X2)  (CKern.h:146-157)
int ref_kern_cross(int ncomp, const int* types, const double* tparams, const double* X, int N, const double* X2, int N2,
                   int D, double* K) {
  REF_TRY
  CCmpndKern* kern = makeKern(ncomp, types, tparams, D);
  if (!kern) return -1;
  CMatrix Xm(N, D, const_cast<double*>(X));
  CMatrix X2m(N2, D, const_cast<double*>(X2));
  CMatrix Km(N, N2);
  kern->compute(Km, Xm, X2m);
  memcpy(K, Km.getVals(), sizeof(double) * (size_t)N * N2);
  // kern (and its clones) leaked on purpose, see makeKern
  return 0;
  REF_CATCH
}

int ref_kern_diag(int ncomp, const int* types, const double* tparams, const double* X, int N, int D, double* d) {
  REF_TRY
  CCmpndKern* kern = makeKern(ncomp, types, tparams, D);
  if (!kern) return -1;
  CMatrix Xm(N, D, const_cast<double*>(X));
  CMatrix dm(N, 1);
  kern->diagCompute(dm, Xm);
  memcpy(d, dm.getVals(), sizeof(double) * N);
  // kern (and its clones) leaked on purpose, see makeKern
  return 0;
  REF_CATCH
}

// g = kern.getGradTransParams(g, X, covGrad, regularise=false)   (CKern.cpp:50-63)
int ref_kern_grad(int ncomp, const int* types, const double* tparams, const double* X, int N, int D,
                  const double* covGrad, double* g) {
  REF_TRY
  CCmpndKern* kern = makeKern(ncomp, types, tparams, D);
  if (!kern) return -1;
  CMatrix Xm(N, D, const_cast<double*>(X));
  CMatrix cg(N, N, const_cast<double*>(covGrad));
  cg.setSymmetric(true);
  CMatrix gm(1, kern->getNumParams());
  kern->getGradTransParams(gm, Xm, cg, false);
  memcpy(g, gm.getVals(), sizeof(double) * kern->getNumParams());
  // kern (and its clones) leaked on purpose, see makeKern
  return 0;
  REF_CATCH
}

// g = kern.getGradTransParams(g, X, X2, covGrad2, regularise=false)   (CKern.cpp:36-49)
int ref_kern_grad2(int ncomp, const int* types, const double* tparams, const double* X, int N, const double* X2,
                   int N2, int D, const double* covGrad, double* g) {
  REF_TRY
  CCmpndKern* kern = makeKern(ncomp, types, tparams, D);
  if (!kern) return -1;
  CMatrix Xm(N, D, const_cast<double*>(X));
  CMatrix X2m(N2, D, const_cast<double*>(X2));
  CMatrix cg(N, N2, const_cast<double*>(covGrad));
  CMatrix gm(1, kern->getNumParams());
  kern->getGradTransParams(gm, Xm, X2m, cg, false);
  memcpy(g, gm.getVals(), sizeof(double) * kern->getNumParams());
  // kern (and its clones) leaked on purpose, see makeKern
  return 0;
  REF_CATCH
}

// gX[i] = kern.getGradX(gX_i, X, i, X2): out is N blocks of N2 x D (CKern.h:68-74)
int ref_kern_gradX(int ncomp, const int* types, const double* tparams, const double* X, int N, const double* X2,
                   int N2, int D, double* out) {
  REF_TRY
  CCmpndKern* kern = makeKern(ncomp, types, tparams, D);
  if (!kern) return -1;
  CMatrix Xm(N, D, const_cast<double*>(X));
  CMatrix X2m(N2, D, const_cast<double*>(X2));
  CMatrix g(N2, D);
  for (int i = 0; i < N; i++) {
    kern->getGradX(g, Xm, i, X2m, false);
    memcpy(out + (size_t)i * N2 * D, g.getVals(), sizeof(double) * (size_t)N2 * D);
  }
  // kern (and its clones) leaked on purpose, see makeKern
  return 0;
  REF_CATCH
}

int ref_kern_diagGradX(int ncomp, const int* types, const double* tparams, const double* X, int N, int D, double* out) {
  REF_TRY
  CCmpndKern* kern = makeKern(ncomp, types, tparams, D);
  if (!kern) return -1;
  CMatrix Xm(N, D, const_cast<double*>(X));
  CMatrix g(N, D);
  kern->getDiagGradX(g, Xm, false);
  memcpy(out, g.getVals(), sizeof(double) * (size_t)N * D);
  // kern (and its clones) leaked on purpose, see makeKern
  return 0;
  REF_CATCH
}

// ---- CMatrix level (lapack.h boundary) ----------------------------------------------------
// U = chol(A) upper (CMatrix.cpp:380-403); returns 1 when MatrixNonPosDef is thrown.
int ref_chol(const double* A, int n, double* U) {
  try {
    CMatrix Am(n, n, const_cast<double*>(A));
    Am.setSymmetric(true);
    Am.chol();
    memcpy(U, Am.getVals(), sizeof(double) * (size_t)n * n);
    return 0;
  } catch (ndlexceptions::MatrixNonPosDef& e) {
    return 1;
  } catch (std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// jitChol (CMatrix.cpp:767-804): U from A with jitter retries; A is mutated like the reference does.
int ref_jitchol(double* A, int n, double* U, double* jitter) {
  REF_TRY
  CMatrix Am(n, n, A);
  Am.setSymmetric(true);
  CMatrix Um(n, n);
  *jitter = Um.jitChol(Am);
  memcpy(U, Um.getVals(), sizeof(double) * (size_t)n * n);
  memcpy(A, Am.getVals(), sizeof(double) * (size_t)n * n);
  return 0;
  REF_CATCH
}
// Ainv = pdinv(U) (CMatrix.cpp:421-432), logdet = logDet(U) (CMatrix.cpp:404-412)
int ref_pdinv(const double* U, int n, double* Ainv, double* logdet) {
  REF_TRY
  CMatrix Um(n, n, const_cast<double*>(U));
  Um.setTriangular(true);
  CMatrix inv(n, n);
  inv.setSymmetric(true);
  inv.pdinv(Um);
  *logdet = logDet(Um);
  memcpy(Ainv, inv.getVals(), sizeof(double) * (size_t)n * n);
  return 0;
  REF_CATCH
}
// B := alpha * op(T^-1) B  etc. (CMatrix.cpp:272-295); T is k x k triangular, B is m x n.
int ref_trsm(double* B, int m, int n, const double* T, double alpha, const char* side, const char* uplo,
             const char* trans, const char* diag) {
  REF_TRY
  int k = (side[0] == 'l' || side[0] == 'L') ? m : n;
  CMatrix Bm(m, n, B);
  CMatrix Tm(k, k, const_cast<double*>(T));
  Tm.setTriangular(true);
  Bm.trsm(Tm, alpha, side, uplo, trans, diag);
  memcpy(B, Bm.getVals(), sizeof(double) * (size_t)m * n);
  return 0;
  REF_CATCH
}
// C := alpha * A A' + beta * C (trans "n") or alpha * A' A + beta * C ("t")  (CMatrix.cpp:297-322)
int ref_syrk(double* C, int n, const double* A, int ar, int ac, double alpha, double beta, const char* uplo,
             const char* trans) {
  REF_TRY
  CMatrix Cm(n, n, C);
  Cm.setSymmetric(true);
  CMatrix Am(ar, ac, const_cast<double*>(A));
  Cm.syrk(Am, alpha, beta, uplo, trans);
  memcpy(C, Cm.getVals(), sizeof(double) * (size_t)n * n);
  return 0;
  REF_CATCH
}
int ref_gemm(double* C, int m, int n, const double* A, int ar, int ac, const double* B, int br, int bc, double alpha,
             double beta, const char* ta, const char* tb) {
  REF_TRY
  CMatrix Cm(m, n, C);
  CMatrix Am(ar, ac, const_cast<double*>(A));
  CMatrix Bm(br, bc, const_cast<double*>(B));
  Cm.gemm(Am, Bm, alpha, beta, ta, tb);
  memcpy(C, Cm.getVals(), sizeof(double) * (size_t)m * n);
  return 0;
  REF_CATCH
}

// ---- CGp level -----------------------------------------------------------------------------
// One logLik+grad evaluation of an FTC CGp built exactly as gp.cpp:379-406 does.
//   out_ll   : CGp::logLikelihood()               (CGp.cpp:913-1013)
//   out_g    : CGp::logLikelihoodGradient(g)      (CGp.cpp:1016-1079), kernel trans-params order
//   timings  : [0] cold logLikelihood (K dirty)  [1] logLikelihoodGradient (K clean)  seconds
//   Xs/Ns    : optional test inputs -> mu, var through CGp::posteriorMeanVar (CGp.cpp:642-663)
//   reps     : evaluate `reps` times (setOptParams between, forcing K dirty) and report the
//              median-free per-eval wall time in timings[2]
int ref_gp_eval(int ncomp, const int* types, const double* tparams, const double* X, const double* y, int N, int D,
                int dout, const double* bias, const double* scale, double* out_ll, double* out_g, const double* Xs,
                int Ns, double* mu, double* var, double* timings, int reps) {
  REF_TRY
  CCmpndKern* kern = makeKern(ncomp, types, tparams, D);
  if (!kern) return -1;
  CMatrix Xm(N, D, const_cast<double*>(X));
  CMatrix ym(N, dout, const_cast<double*>(y));
  CGaussianNoise noise(&ym);
  CMatrix nbias(1, dout, 0.0);  // gp.cpp:380 passes 0.0 (implicit 1x1 CMatrix): only valid for dout==1
  noise.setBias(nbias);
  CGp model(kern, &noise, &Xm, CGp::FTC, 0, 0);
  model.setBetaVal(1);
  CMatrix sc(1, dout, const_cast<double*>(scale));
  CMatrix bi(1, dout, const_cast<double*>(bias));
  model.setScale(sc);
  model.setBias(bi);
  model.updateM();
  unsigned int P = model.getOptNumParams();
  CMatrix g(1, P);
  CMatrix params(1, P);
  model.getOptParams(params);
  double t0 = now();
  double ll = model.logLikelihood();
  double t1 = now();
  model.logLikelihoodGradient(g);
  double t2 = now();
  *out_ll = ll;
  for (unsigned int i = 0; i < P; i++) out_g[i] = g.getVal(0, i);
  if (timings) {
    timings[0] = t1 - t0;
    timings[1] = t2 - t1;
    timings[2] = t2 - t0;
  }
  if (reps > 1 && timings) {
    double ta = now();
    for (int r = 0; r < reps; r++) {
      model.setOptParams(params);  // marks K dirty (CGp.cpp:387-389)
      model.logLikelihoodGradient(g);
    }
    timings[2] = (now() - ta) / reps;
  }
  if (Xs && Ns > 0) {
    CMatrix Xsm(Ns, D, const_cast<double*>(Xs));
    CMatrix mum(Ns, dout), varm(Ns, dout);
    model.posteriorMeanVar(mum, varm, Xsm);
    memcpy(mu, mum.getVals(), sizeof(double) * (size_t)Ns * dout);
    memcpy(var, varm.getVals(), sizeof(double) * (size_t)Ns * dout);
  }
  // kern (and its clones) leaked on purpose, see makeKern
  return (int)P;
  REF_CATCH
}

// Phase timings of the reference path on this host: kern.compute, chol, pdinv, trans.
int ref_phase_times(int ncomp, const int* types, const double* tparams, const double* X, int N, int D, double* t) {
  REF_TRY
  CCmpndKern* kern = makeKern(ncomp, types, tparams, D);
  if (!kern) return -1;
  CMatrix Xm(N, D, const_cast<double*>(X));
  CMatrix K(N, N);
  double t0 = now();
  kern->compute(K, Xm);
  double t1 = now();
  CMatrix U(N, N);
  U.deepCopy(K);
  U.setSymmetric(true);
  double t2 = now();
  U.potrf("U");
  double t3 = now();
  CMatrix inv(N, N);
  inv.setSymmetric(true);
  U.setTriangular(true);
  // zero strict lower like chol() does so pdinv sees a clean factor
  for (int j = 0; j < N; j++)
    for (int i = j + 1; i < N; i++) U.setVal(0.0, i, j);
  double t4 = now();
  inv.pdinv(U);
  double t5 = now();
  U.trans();
  double t6 = now();
  t[0] = t1 - t0;  // kern.compute
  t[1] = t3 - t2;  // dpotrf_
  t[2] = t5 - t4;  // pdinv (dpotri_ + mirror)
  t[3] = t6 - t5;  // trans (Alg. 513)
  // kern (and its clones) leaked on purpose, see makeKern
  return 0;
  REF_CATCH
}

// GP-LVM: one CGplvm logLikelihood + gradient at latent X (CGplvm.cpp:493-716), noise = CScaleNoise as
// gplvm.cpp:504-522 builds it.  Parameter / gradient layout follows CGplvm::getOptParams
// (CGplvm.cpp:257-290): [kernel trans-params][X col-major].  out_m receives the N x dout matrix m
// ((y-bias)/scale, CNoise.cpp:710-721) so the device path can be fed identical targets.
int ref_gplvm_eval(int ncomp, const int* types, const double* tparams, const double* Xlat, const double* Y, int N,
                   int q, int dout, double* out_ll, double* out_g, double* out_m) {
  REF_TRY
  CCmpndKern* kern = makeKern(ncomp, types, tparams, q);
  if (!kern) return -1;
  CMatrix Ym(N, dout, const_cast<double*>(Y));
  CScaleNoise noise(&Ym);
  CGplvm model(kern, &noise, q, 0);
  unsigned int P = model.getOptNumParams();
  unsigned int nk = kern->getNumParams();
  CMatrix params(1, P);
  for (unsigned int i = 0; i < nk; i++) params.setVal(tparams[i], i);
  for (size_t i = 0; i < (size_t)N * q; i++) params.setVal(Xlat[i], nk + i);
  model.setOptParams(params);
  CMatrix g(1, P);
  *out_ll = model.logLikelihood();
  model.logLikelihoodGradient(g);
  for (unsigned int i = 0; i < P; i++) out_g[i] = g.getVal(0, i);
  if (out_m) memcpy(out_m, model.m.getVals(), sizeof(double) * (size_t)N * dout);
  // kern (and its clones) leaked on purpose, see makeKern
  return (int)P;
  REF_CATCH
}

// PCA initialisation of the latent space, as the CGplvm constructor leaves it (CGplvm.cpp:157-222).
int ref_gplvm_initX(const double* Y, int N, int q, int dout, double* Xout) {
  REF_TRY
  int types[3] = {K_RBF, K_BIAS, K_WHITE};
  double tp[4] = {0, 0, -2, -2};
  CCmpndKern* kern = makeKern(3, types, tp, q);
  CMatrix Ym(N, dout, const_cast<double*>(Y));
  CScaleNoise noise(&Ym);
  CGplvm model(kern, &noise, q, 0);
  memcpy(Xout, model.pX->getVals(), sizeof(double) * (size_t)N * q);
  // kern (and its clones) leaked on purpose, see makeKern
  return 0;
  REF_CATCH
}

// The reference's SVM-light reader (CClctrl::readSvmlDataFile, CClctrl.cpp:55-171).  X == NULL: sizes only.
int ref_svml_read(const char* path, int* n, int* d, double* X, double* y) {
  REF_TRY
  struct Ctl : public CClctrl {  // CClctrl is abstract only in its help texts (CClctrl.h:50-51)
    Ctl(int c, char** v) : CClctrl(c, v) {}
    void helpInfo() {}
    void helpHeader() {}
  };
  char arg0[] = "ref";
  char* argv[] = {arg0, 0};
  Ctl ctl(1, argv);
  CMatrix Xm, ym;
  ctl.readSvmlDataFile(Xm, ym, std::string(path));
  *n = (int)Xm.getRows();
  *d = (int)Xm.getCols();
  if (X) memcpy(X, Xm.getVals(), sizeof(double) * (size_t)Xm.getRows() * Xm.getCols());
  if (y) memcpy(y, ym.getVals(), sizeof(double) * (size_t)ym.getRows());
  return 0;
  REF_CATCH
}

// CGp::optimise (SCG, the default optimiser of `gp learn`, gp.cpp:404) for `iters` iterations from the given
// transformed parameters; returns the transformed parameters and the final log-likelihood.
int ref_gp_optimise(int ncomp, const int* types, const double* tparams, const double* X, const double* y, int N, int D,
                    int dout, const double* bias, const double* scale, int iters, double* out_tparams, double* out_ll) {
  REF_TRY
  CCmpndKern* kern = makeKern(ncomp, types, tparams, D);
  if (!kern) return -1;
  CMatrix Xm(N, D, const_cast<double*>(X));
  CMatrix ym(N, dout, const_cast<double*>(y));
  CGaussianNoise noise(&ym);
  CMatrix nbias(1, dout, 0.0);
  noise.setBias(nbias);
  CGp model(kern, &noise, &Xm, CGp::FTC, 0, 0);  // verbosity 0: no display / checkGradients
  model.setBetaVal(1);
  CMatrix sc(1, dout, const_cast<double*>(scale));
  CMatrix bi(1, dout, const_cast<double*>(bias));
  model.setScale(sc);
  model.setBias(bi);
  model.updateM();
  model.setDefaultOptimiser(CGp::SCG);  // gp.cpp:393-394
  model.optimise(iters);                // setMaxIters + runDefaultOptimiser (CGp.cpp:1537-1553)
  unsigned int P = kern->getNumParams();
  CMatrix tp(1, P);
  kern->getTransParams(tp);
  for (unsigned int i = 0; i < P; i++) out_tparams[i] = tp.getVal(0, i);
  *out_ll = model.logLikelihood();
  return (int)P;
  REF_CATCH
}

}  // extern "C"
