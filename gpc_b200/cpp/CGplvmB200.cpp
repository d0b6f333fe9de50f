// CGplvmB200.cpp -- see CGplvmB200.h.  Builds with the reference's own flags (-std=gnu++98 -D_LINUX).
#include "CGplvmB200.h"
#include <cstdlib>
#include <cmath>

CGplvmB200::CGplvmB200() : CGplvm() { init(); }
CGplvmB200::CGplvmB200(CKern* kernel, CScaleNoise* nois, const int latDim, const int verbos)
    : CGplvm(kernel, nois, latDim, verbos)
{
  init();
}
CGplvmB200::CGplvmB200(CKern* kernel, CKern* dynKernel, CScaleNoise* nois, const int latDim, const int verbos)
    : CGplvm(kernel, dynKernel, nois, latDim, verbos)
{
  init();
}
CGplvmB200::CGplvmB200(CKern* kernel, CMatrix* backKernel, CScaleNoise* nois, const int latDim, const int verbos)
    : CGplvm(kernel, backKernel, nois, latDim, verbos)
{
  init();
}
CGplvmB200::CGplvmB200(CKern* kernel, CKern* dynKernel, CMatrix* backKernel, CScaleNoise* nois, const int latDim,
                       const int verbos)
    : CGplvm(kernel, dynKernel, backKernel, nois, latDim, verbos)
{
  init();
}
CGplvmB200::~CGplvmB200()
{
  if(dev)
    gpc_ctx_destroy(dev);
}
void CGplvmB200::init()
{
  dev = 0;
  devN = 0;
  devD = devd = 0;
  const char* e = getenv("GPC_DEVICE");
  device = e ? atoi(e) : 0;
  fresh = false;
  evalOut[0] = evalOut[1] = evalOut[2] = 0.0;
  nEvals = 0;
}
void CGplvmB200::setDevice(int d)
{
  if(d != device && dev)
  {
    gpc_ctx_destroy(dev);
    dev = 0;
  }
  device = d;
  fresh = false;
}
void CGplvmB200::fail(int rc) const
{
  fresh = false;
  if(rc > 0)
    throw ndlexceptions::MatrixNonPosDef();
  throw ndlexceptions::Error(std::string("gpc_b200: ") + gpc_last_error());
}

bool CGplvmB200::onDevice() const
{
  if(isDynamicModelLearnt() || isBackConstrained() || isSparseApproximation() || !pkern || !pX)
    return false;
  return bridge.sync(pkern, pX->getCols());
}

// everything the cached evaluation was computed from: kernel parameters, the latent positions *pX and the targets m.
// O(N (q + d)) per call against O(N^3) per evaluation; catches changes that do not go through setOptParams / updateX
// (initXpca / initXrand after a first evaluation, pnoise->setScale / setBias / setTarget + updateSites)
static void gplvmKey(std::vector<double>& k, const std::vector<double>& p, const CMatrix& X, const CMatrix& m)
{
  k = p;
  k.insert(k.end(), X.getVals(), X.getVals() + (size_t)X.getRows() * X.getCols());
  k.insert(k.end(), m.getVals(), m.getVals() + (size_t)m.getRows() * m.getCols());
}

void CGplvmB200::ensureEvaluated() const
{
  std::vector<double> now;
  gplvmKey(now, bridge.naturalParams(), *pX, m);
  if(fresh && key == now)
    return;
  int64_t N = pX->getRows();
  int D = (int)pX->getCols(), d = (int)m.getCols();
  DIMENSIONMATCH(m.getRows() == pX->getRows());
  if(!dev || N != devN || D != devD || d != devd)
  {
    if(dev)
      gpc_ctx_destroy(dev);
    dev = 0;
    int rc = gpc_ctx_create(&dev, device, N, D, d);
    if(rc)
      fail(rc);
    devN = N;
    devD = D;
    devd = d;
  }
  int rc = gpc_set_X(dev, pX->getVals(), N, D, N);
  if(rc)
    fail(rc);
  rc = gpc_set_M(dev, m.getVals(), N, d, N); // m is kept by the noise model (CScaleNoise::updateSites, CNoise.cpp:710-721)
  if(rc)
    fail(rc);
  gNat.assign(bridge.getNumParams(), 0.0);
  gLatent.assign((size_t)N * D, 0.0);
  rc = gpc_eval(dev, bridge.comps(), bridge.numComps(), 1, evalOut, &gNat[0], &gLatent[0]);
  if(rc)
    fail(rc);
  nEvals++;
  if(evalOut[2] > 0.0)
  {
    // CGplvm::_updateInvK factorises with plain chol(), no jitter retry (CGplvm.cpp:437-445): where the library had to
    // add jitter the reference throws
    fresh = false;
    throw ndlexceptions::MatrixNonPosDef();
  }
  key.swap(now);
  fresh = true;
}

double CGplvmB200::logLikelihood() const
{
  if(!onDevice())
    return CGplvm::logLikelihood();
  ensureEvaluated();
  double L = evalOut[1] + (double)getNumProcesses() * evalOut[0]; // CGplvm.cpp:497-507
  if(isLatentRegularised())
    for(int j = 0; j < getLatentDim(); j++)
      L += pX->norm2Col(j); // Gaussian prior over the latent positions (CGplvm.cpp:528-538)
  if(isInputScaleLearnt())
    for(int j = 0; j < getNumProcesses(); j++)
      L += 2 * log(fabs(pnoise->getScale(j))); // CGplvm.cpp:540-545
  L *= -0.5;
  L += pkern->priorLogProb();
  return L;
}

double CGplvmB200::logLikelihoodGradient(CMatrix& g) const
{
  if(!onDevice())
    return CGplvm::logLikelihoodGradient(g);
  ensureEvaluated();
  unsigned int P = bridge.getNumParams();
  unsigned int N = getNumData();
  int q = getLatentDim();
  DIMENSIONMATCH(g.getRows() == 1 && g.getCols() == getOptNumParams());
  g.zeros();
  std::vector<double> gk(gNat);
  bridge.finishGradient(pkern, &gk[0]);
  for(unsigned int i = 0; i < P; i++)
    g.setVal(gk[i], 0, i);
  // dL/dX, column-major behind the kernel parameters (CGplvm.cpp:594-603), minus X for the latent prior (:672-681)
  for(int k = 0; k < q; k++)
    for(unsigned int i = 0; i < N; i++)
    {
      double v = gLatent[i + (size_t)N * k];
      if(isLatentRegularised())
        v -= pX->getVal(i, k);
      g.setVal(v, 0, P + i + N * k);
    }
  if(isInputScaleLearnt())
  {
    // (m_j' K^-1 m_j - 1)/scale_j behind the latent block (CGplvm.cpp:700-711)
    int d = getNumProcesses();
    CMatrix A(N, d);
    int rc = gpc_download(dev, GPC_MAT_ALPHA, A.getVals(), N);
    if(rc)
      fail(rc);
    for(int j = 0; j < d; j++)
      g.setVal(1 / pnoise->getScale(j) * (A.dotColCol(j, m, j) - 1), 0, P + N * q + j);
  }
  return logLikelihood();
}

void CGplvmB200::setOptParams(const CMatrix& param)
{
  CGplvm::setOptParams(param); // kernel parameters, *pX in place, scales (CGplvm.cpp:292-330)
  fresh = false;
}
void CGplvmB200::updateX()
{
  CGplvm::updateX();
  fresh = false;
}

void CGplvmB200::posteriorMeanVar(CMatrix& mu, CMatrix& varSigma, const CMatrix& Xin) const
{
  if(!onDevice())
  {
    CGplvm::posteriorMeanVar(mu, varSigma, Xin);
    return;
  }
  DIMENSIONMATCH(mu.getCols() == (unsigned int)getNumProcesses());
  DIMENSIONMATCH(varSigma.getCols() == (unsigned int)getNumProcesses());
  DIMENSIONMATCH(mu.getRows() == Xin.getRows() && varSigma.getRows() == Xin.getRows());
  ensureEvaluated();
  double quad = 0.0;
  int rc = gpc_solve_alpha(dev, &quad);
  if(rc)
    fail(rc);
  // mean K*' K^-1 m and variance k** - |L^-1 k*|^2, both in the space of m (CGplvm.cpp:340-362)
  rc = gpc_posterior(dev, bridge.comps(), bridge.numComps(), Xin.getVals(), Xin.getRows(), Xin.getRows(), mu.getVals(),
                     varSigma.getVals());
  if(rc)
    fail(rc);
}

CGplvmB200* readGplvmB200FromStream(istream& in)
{
  CGplvmB200* pmodel = new CGplvmB200();
  pmodel->fromStream(in);
  return pmodel;
}
CGplvmB200* readGplvmB200FromFile(const string modelFileName, const int verbosity)
{
  if(verbosity > 0)
    cout << "Loading model file." << endl;
  ifstream in(modelFileName.c_str());
  if(!in.is_open())
    throw ndlexceptions::FileReadError(modelFileName);
  CGplvmB200* pmodel;
  try
  {
    pmodel = readGplvmB200FromStream(in);
  }
  catch(ndlexceptions::FileFormatError& err)
  {
    throw ndlexceptions::FileFormatError(modelFileName);
  }
  if(verbosity > 0)
    cout << "... done." << endl;
  in.close();
  return pmodel;
}
