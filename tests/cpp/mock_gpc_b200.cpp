// mock_gpc_b200.cpp -- TEST DOUBLE (tests/test_cpp_host_cpu.py), never shipped and never loaded by the product.
//
// A stand-in for libgpc_b200.so that implements the subset of the C ABI the C++ host classes call (include/gpc_b200.h:
// gpc_ctx_*, gpc_set_X/M, gpc_eval, gpc_kern_build, gpc_jitchol, gpc_solve_alpha, gpc_posterior, gpc_download) on the
// host with the REFERENCE's own classes (CKern / CMatrix from /root/reference, linked in by oracle/build_ref.sh).  It
// exists so that the device-path logic of gpc_b200/cpp -- uploads, result caching and invalidation, gradient assembly,
// scale / bias handling, error mapping -- is exercised on a machine without a GPU (and under AddressSanitizer), the same
// way tests/test_dist_cpu.py drives the multi-GPU schedule with numpy kernels.  Placed in front of the real library with
// LD_LIBRARY_PATH by the tests only; it says so on stderr when a context is created.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "CKern.h"
#include "CMatrix.h"
#include "gpc_b200.h"

struct gpc_ctx
{
  int64_t N;
  int D, d;
  CMatrix X, M, K, L, Kinv, Alpha; // L lower
  bool haveX, haveM, haveL, haveAlpha;
  long evals;
};

static std::string g_err;
static long g_total_evals = 0;

static CKern* leafOf(const gpc_kcomp& c, int D)
{
  CKern* k = 0;
  switch(c.type)
  {
  case GPC_KERN_WHITE: k = new CWhiteKern(D); break;
  case GPC_KERN_BIAS: k = new CBiasKern(D); break;
  case GPC_KERN_RBF: k = new CRbfKern(D); break;
  case GPC_KERN_RBFARD: k = new CRbfardKern(D); break;
  case GPC_KERN_MATERN32: k = new CMatern32Kern(D); break;
  case GPC_KERN_MATERN52: k = new CMatern52Kern(D); break;
  case GPC_KERN_LIN: k = new CLinKern(D); break;
  case GPC_KERN_POLY:
  {
    CPolyKern* p = new CPolyKern(D);
    p->setDegree(c.degree);
    k = p;
    break;
  }
  default: return 0;
  }
  if((int)k->getNumParams() != c.nparams)
    return 0;
  for(int i = 0; i < c.nparams; i++)
    k->setParam(c.params[i], i);
  return k;
}
// the component objects are leaked on purpose (the reference's ARD clones share their scale buffer with the original)
static CCmpndKern* kernOf(const gpc_kcomp* comps, int ncomp, int D)
{
  CMatrix Xtmp(1, D);
  CCmpndKern* kern = new CCmpndKern(Xtmp);
  for(int c = 0; c < ncomp; c++)
  {
    CKern* k = leafOf(comps[c], D);
    if(!k)
      return 0;
    kern->addKern(k);
  }
  return kern;
}

static void buildK(const CKern* kern, const CMatrix& X, CMatrix& K)
{
  unsigned int N = X.getRows();
  K.resize(N, N);
  for(unsigned int i = 0; i < N; i++) // CGp::_updateK (CGp.cpp:693-712)
  {
    K.setVal(kern->diagComputeElement(X, i), i, i);
    for(unsigned int j = 0; j < i; j++)
    {
      double v = kern->computeElement(X, i, X, j);
      K.setVal(v, i, j);
      K.setVal(v, j, i);
    }
  }
  K.setSymmetric(true);
}

// the reference's jitter schedule (CMatrix.cpp:767-804); *added = the total jitter actually added (0 if none)
static int factor(gpc_ctx* c, double* added, double* logdet)
{
  *added = 0.0;
  double jitter = 1e-6 * c->K.trace() / (double)c->K.getRows();
  for(int tries = 0; tries < 20; tries++)
  {
    try
    {
      c->L.deepCopy(c->K);
      c->L.chol(); // upper
      *logdet = logDet(c->L);
      c->L.trans();
      c->haveL = true;
      return 0;
    }
    catch(ndlexceptions::MatrixNonPosDef&)
    {
      c->K.addDiag(jitter);
      *added += jitter;
      jitter *= 10;
      if(jitter > 10)
        break;
    }
  }
  g_err = "mock: kernel matrix is non positive definite after jitter retries";
  return 1;
}

extern "C" {

const char* gpc_last_error(void) { return g_err.c_str(); }
int gpc_device_count(void) { return 1; }
int gpc_kern_nparams(int type, int D)
{
  switch(type)
  {
  case GPC_KERN_WHITE: case GPC_KERN_BIAS: case GPC_KERN_LIN: return 1;
  case GPC_KERN_RBF: case GPC_KERN_MATERN32: case GPC_KERN_MATERN52: return 2;
  case GPC_KERN_RBFARD: return 2 + D;
  case GPC_KERN_POLY: return 3;
  }
  return -1;
}

int gpc_ctx_create(gpc_ctx** out, int device, int64_t Nmax, int Dmax, int dout_max)
{
  fprintf(stderr, "mock_gpc_b200: host test double in use (N<=%lld)\n", (long long)Nmax);
  gpc_ctx* c = new gpc_ctx();
  c->N = 0;
  c->D = c->d = 0;
  c->haveX = c->haveM = c->haveL = c->haveAlpha = false;
  c->evals = 0;
  *out = c;
  return GPC_OK;
}
int gpc_ctx_destroy(gpc_ctx* c)
{
  delete c;
  return GPC_OK;
}
int gpc_ctx_sync(gpc_ctx*) { return GPC_OK; }
int64_t gpc_ctx_launch_count(gpc_ctx* c) { return c->evals; }

int gpc_set_X(gpc_ctx* c, const double* X, int64_t N, int D, int64_t ldx)
{
  c->X.resize(N, D);
  for(int j = 0; j < D; j++)
    for(int64_t i = 0; i < N; i++)
      c->X.setVal(X[i + j * ldx], i, j);
  c->N = N;
  c->D = D;
  c->haveX = true;
  c->haveL = c->haveAlpha = false;
  return GPC_OK;
}
int gpc_set_M(gpc_ctx* c, const double* M, int64_t N, int d, int64_t ldm)
{
  c->M.resize(N, d);
  for(int j = 0; j < d; j++)
    for(int64_t i = 0; i < N; i++)
      c->M.setVal(M[i + j * ldm], i, j);
  c->d = d;
  c->haveM = true;
  c->haveAlpha = false;
  return GPC_OK;
}

int gpc_kern_build(gpc_ctx* c, const gpc_kcomp* comps, int ncomp)
{
  CCmpndKern* kern = kernOf(comps, ncomp, c->D);
  if(!kern || !c->haveX)
  {
    g_err = "mock: bad kernel specification or no X";
    return GPC_ERR_ARG;
  }
  buildK(kern, c->X, c->K);
  c->haveL = c->haveAlpha = false;
  return GPC_OK;
}
int gpc_kern_cross(gpc_ctx* c, const gpc_kcomp* comps, int ncomp, const double* Xs, int64_t Ns, int64_t ldxs, double* Ks,
                   int64_t ldk)
{
  CCmpndKern* kern = kernOf(comps, ncomp, c->D);
  if(!kern || !c->haveX)
  {
    g_err = "mock: bad kernel specification or no X";
    return GPC_ERR_ARG;
  }
  CMatrix X2(Ns, c->D);
  for(int j = 0; j < c->D; j++)
    for(int64_t i = 0; i < Ns; i++)
      X2.setVal(Xs[i + j * ldxs], i, j);
  for(int64_t j = 0; j < Ns; j++) // computeElement for every pair (CKern.h:146-157)
    for(int64_t i = 0; i < c->N; i++)
      Ks[i + j * ldk] = kern->computeElement(c->X, i, X2, j);
  delete kern;
  return GPC_OK;
}
int gpc_jitchol(gpc_ctx* c, int max_tries, double* jitter, double* logdet)
{
  double added = 0.0, ld = 0.0;
  int rc = factor(c, &added, &ld);
  if(jitter)
    *jitter = added;
  if(logdet)
    *logdet = ld;
  return rc;
}
int gpc_solve_alpha(gpc_ctx* c, double* quad)
{
  if(!c->haveL || !c->haveM)
  {
    g_err = "mock: state: gpc_solve_alpha needs L and m";
    return GPC_ERR_STATE;
  }
  c->Alpha.deepCopy(c->M); // CGp::updateAlpha (CGp.cpp:469-484)
  c->Alpha.trsm(c->L, 1.0, "L", "L", "N", "N");
  c->Alpha.trsm(c->L, 1.0, "L", "L", "T", "N");
  c->haveAlpha = true;
  if(quad)
  {
    *quad = 0.0;
    for(int j = 0; j < c->d; j++)
      *quad += c->Alpha.dotColCol(j, c->M, j);
  }
  return GPC_OK;
}

int gpc_eval(gpc_ctx* c, const gpc_kcomp* comps, int ncomp, int flags, double* out, double* gparams, double* gX)
{
  if(!c->haveX || !c->haveM)
  {
    g_err = "mock: state: gpc_eval needs X and m";
    return GPC_ERR_STATE;
  }
  CCmpndKern* kern = kernOf(comps, ncomp, c->D);
  if(!kern)
  {
    g_err = "mock: bad kernel specification";
    return GPC_ERR_ARG;
  }
  buildK(kern, c->X, c->K);
  double added = 0.0, logdet = 0.0;
  int rc = factor(c, &added, &logdet);
  if(rc)
    return rc;
  CMatrix U(c->L);
  U.trans();
  c->Kinv.resize(c->N, c->N);
  c->Kinv.setSymmetric(true);
  c->Kinv.pdinv(U);
  unsigned int N = c->N, d = c->d, P = kern->getNumParams(), D = c->D;
  c->Alpha.resize(N, d);
  double quad = 0.0;
  for(unsigned int j = 0; j < d; j++)
  {
    c->Alpha.symvColCol(j, c->Kinv, c->M, j, 1.0, 0.0, "u");
    quad += c->Alpha.dotColCol(j, c->M, j);
  }
  c->haveAlpha = true;
  if(out)
  {
    out[0] = logdet;
    out[1] = quad;
    out[2] = added;
  }
  bool wantX = (flags & 1) && gX;
  std::vector<double> g(P, 0.0);
  CMatrix gXacc(N, D);
  gXacc.zeros();
  std::vector<CMatrix*> gKX;
  CMatrix dgKX(N, D);
  if(wantX)
  {
    for(unsigned int i = 0; i < N; i++)
      gKX.push_back(new CMatrix(N, D));
    static_cast<const CKern*>(kern)->getGradX(gKX, c->X, c->X); // CGplvm.cpp:569-581 (the vector form lives in CKern)
    kern->getDiagGradX(dgKX, c->X);
    for(unsigned int i = 0; i < N; i++)
    {
      gKX[i]->scale(2.0);
      for(unsigned int k = 0; k < D; k++)
        gKX[i]->setVal(dgKX.getVal(i, k), i, k);
    }
  }
  CMatrix covGrad(N, N), tmpG(1, P), a(N, 1);
  for(unsigned int j = 0; j < d; j++)
  {
    // covGrad = -1/2 (K^-1 - alpha_j alpha_j') (CGp::updateCovGradient, CGp.cpp:666-679)
    a.copyColCol(0, c->Alpha, j);
    covGrad.deepCopy(c->Kinv);
    covGrad.syr(a, -1.0, "u");
    covGrad.scale(-0.5);
    for(unsigned int p = 0; p < N; p++) // syr touched the upper triangle only
      for(unsigned int q = 0; q < p; q++)
        covGrad.setVal(covGrad.getVal(q, p), p, q);
    covGrad.setSymmetric(true);
    kern->getGradParams(tmpG, c->X, covGrad, false);
    for(unsigned int p = 0; p < P; p++)
      g[p] += tmpG.getVal(p);
    if(wantX)
      for(unsigned int i = 0; i < N; i++)
        for(unsigned int k = 0; k < D; k++)
          gXacc.addVal(gKX[i]->dotColCol(k, covGrad, i), i, k); // CGplvm.cpp:594-603
  }
  if(gparams)
    for(unsigned int p = 0; p < P; p++)
      gparams[p] = g[p];
  if(wantX)
  {
    for(unsigned int k = 0; k < D; k++)
      for(unsigned int i = 0; i < N; i++)
        gX[i + (size_t)k * N] = gXacc.getVal(i, k);
    for(unsigned int i = 0; i < N; i++)
      delete gKX[i];
  }
  c->evals++;
  g_total_evals++;
  return GPC_OK;
}

int gpc_posterior(gpc_ctx* c, const gpc_kcomp* comps, int ncomp, const double* Xs, int64_t Ns, int64_t ldxs, double* mu,
                  double* var)
{
  if(!c->haveL || !c->haveAlpha)
  {
    g_err = "mock: state: gpc_posterior needs L and alpha";
    return GPC_ERR_STATE;
  }
  CCmpndKern* kern = kernOf(comps, ncomp, c->D);
  if(!kern)
  {
    g_err = "mock: bad kernel specification";
    return GPC_ERR_ARG;
  }
  CMatrix Xt(Ns, c->D);
  for(int j = 0; j < c->D; j++)
    for(int64_t i = 0; i < Ns; i++)
      Xt.setVal(Xs[i + j * ldxs], i, j);
  CMatrix kX(c->N, Ns);
  kern->compute(kX, c->X, Xt); // CGp::_testComputeKx (CGp.cpp:535-547)
  for(int64_t i = 0; i < Ns; i++)
    for(int j = 0; j < c->d; j++)
      mu[i + (size_t)j * Ns] = c->Alpha.dotColCol(j, kX, i); // CGp.cpp:553-559
  if(var)
  {
    kX.trsm(c->L, 1.0, "L", "L", "N", "N"); // CGp.cpp:603-611
    for(int64_t i = 0; i < Ns; i++)
    {
      double v = kern->diagComputeElement(Xt, i) - kX.norm2Col(i);
      for(int j = 0; j < c->d; j++)
        var[i + (size_t)j * Ns] = v;
    }
  }
  return GPC_OK;
}

// the sparse-approximation entry points are not modelled by this test double: CGpB200 reports the error like any other
// library failure (the device path of the sparse models is covered by the GPU tests)
int gpc_sparse_create(gpc_sparse** out, int, int, int64_t, int, int, int)
{
  if(out)
    *out = 0;
  g_err = "mock: sparse approximations are not modelled by the test double";
  return GPC_ERR_CUDA;
}
int gpc_sparse_destroy(gpc_sparse*) { return GPC_OK; }
int gpc_sparse_set_data(gpc_sparse*, const double*, int64_t, const double*, int64_t) { return GPC_ERR_CUDA; }
int gpc_sparse_eval(gpc_sparse*, const gpc_kcomp*, int, const double*, int64_t, double, double*, double*, double*, double*)
{
  return GPC_ERR_CUDA;
}
int gpc_sparse_posterior(gpc_sparse*, const gpc_kcomp*, int, const double*, int64_t, int64_t, double*, double*)
{
  return GPC_ERR_CUDA;
}

int gpc_download(gpc_ctx* c, int which, double* dst, int64_t ld)
{
  const CMatrix* src = which == GPC_MAT_K ? &c->K : which == GPC_MAT_L ? &c->L : which == GPC_MAT_KINV ? &c->Kinv :
                       which == GPC_MAT_ALPHA ? &c->Alpha : &c->M;
  for(unsigned int j = 0; j < src->getCols(); j++)
    for(unsigned int i = 0; i < src->getRows(); i++)
      dst[i + (size_t)j * ld] = src->getVal(i, j);
  return GPC_OK;
}

} // extern "C"
