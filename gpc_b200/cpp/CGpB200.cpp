// CGpB200.cpp -- see CGpB200.h.  Builds with the reference's own flags (-std=gnu++98 -D_LINUX) against its headers.
#include "CGpB200.h"
#include <cstdlib>
#include <cmath>

namespace
{
// CGp keeps its noise model private (CGp.h:403) and offers no accessor; out() needs it (CGp.cpp:445-460).  An explicit
// template instantiation may name a private member (the access rules do not apply there), which hands us the member
// pointer without touching the reference header.  A maintainer would add `CNoise* getNoise() const` instead.
template <typename Tag, typename Tag::type Member> struct PrivateMember
{
  friend typename Tag::type memberOf(Tag) { return Member; }
};
struct CGpNoiseTag
{
  typedef CNoise* CGp::*type;
  friend type memberOf(CGpNoiseTag);
};
template struct PrivateMember<CGpNoiseTag, &CGp::pnoise>;
// the scaled and biased targets m (CGp.h:364): kept by CGp::updateM under the reference's own MupToDate rules, which
// the device path has to follow exactly (see upload())
struct CGpMTag
{
  typedef CMatrix CGp::*type;
  friend type memberOf(CGpMTag);
};
template struct PrivateMember<CGpMTag, &CGp::m>;

int defaultDevice()
{
  const char* e = getenv("GPC_DEVICE");
  return e ? atoi(e) : 0;
}
} // namespace

CGpB200::CGpB200() : CGp() { init(); }
CGpB200::CGpB200(CKern* kernel, CNoise* nois, CMatrix* Xin, int approxType, unsigned int actSetSize, int verbos)
    : CGp(kernel, nois, Xin, approxType, actSetSize, verbos)
{
  init();
}
CGpB200::CGpB200(unsigned int q, unsigned int d, CMatrix* Xin, CMatrix* yin, CKern* kernel, CNoise* nois, int approxType,
                 unsigned int actSetSize, int verbos)
    : CGp(q, d, Xin, yin, kernel, nois, approxType, actSetSize, verbos)
{
  init();
}
CGpB200::~CGpB200()
{
  if(dev)
    gpc_ctx_destroy(dev);
  if(sdev)
    gpc_sparse_destroy(sdev);
}
void CGpB200::init()
{
  dev = 0;
  devN = 0;
  devD = devd = 0;
  device = defaultDevice();
  state = STALE;
  keyX = keyY = 0;
  evalOut[0] = evalOut[1] = evalOut[2] = 0.0;
  nEvals = 0;
  sdev = 0;
  sN = 0;
  sM = sD = sd = sApprox = 0;
  sstate = STALE;
  sgBeta = 0.0;
  for(int i = 0; i < 6; i++)
    spOut[i] = 0.0;
}
void CGpB200::setDevice(int d)
{
  if(d != device && dev)
  {
    gpc_ctx_destroy(dev);
    dev = 0;
  }
  device = d;
  state = STALE;
}

void CGpB200::fail(int rc) const
{
  state = STALE;
  sstate = STALE;
  if(rc > 0)
    throw ndlexceptions::MatrixNonPosDef(); // what CMatrix::potrf / jitChol throw (CMatrix.cpp:378, 791, 801)
  throw ndlexceptions::Error(std::string("gpc_b200: ") + gpc_last_error());
}

bool CGpB200::onDevice() const
{
  if(isSparseApproximation() || getApproximationType() != FTC || isOptimiseX() || isBackConstrained() || !isSpherical())
    return false;
  if(!pX || !py || !getKernel())
    return false;
  if(isOutputScaleLearnt() && getOutputDim() > 1) // the reference throws here (see logLikelihoodGradient): keep that
    return false;
  return bridge.sync(getKernel(), pX->getCols());
}

// call after bridge.sync(): compares what the cached evaluation was computed from
bool CGpB200::sameInputs() const
{
  if(keyX != pX || keyY != py || !dev || devN != (int64_t)pX->getRows())
    return false;
  const std::vector<double>& p = bridge.naturalParams();
  const CMatrix& mm = this->*memberOf(CGpMTag());
  unsigned int d = getOutputDim();
  size_t nm = (size_t)mm.getRows() * mm.getCols(), nx = (size_t)pX->getRows() * pX->getCols();
  if(key.size() != p.size() + 2 * d + nm + nx)
    return false;
  for(size_t i = 0; i < p.size(); i++)
    if(key[i] != p[i])
      return false;
  for(unsigned int j = 0; j < d; j++)
    if(key[p.size() + j] != getScaleVal(j) || key[p.size() + d + j] != getBiasVal(j))
      return false;
  const double* mv = mm.getVals();
  for(size_t i = 0; i < nm; i++)
    if(key[p.size() + 2 * d + i] != mv[i])
      return false;
  const double* xv = pX->getVals(); // *pX changed in place without updateX() (O(N D) against O(N^3) per evaluation)
  for(size_t i = 0; i < nx; i++)
    if(key[p.size() + 2 * d + nm + i] != xv[i])
      return false;
  return true;
}
void CGpB200::snapshotInputs() const
{
  key = bridge.naturalParams();
  for(unsigned int j = 0; j < getOutputDim(); j++)
    key.push_back(getScaleVal(j));
  for(unsigned int j = 0; j < getOutputDim(); j++)
    key.push_back(getBiasVal(j));
  const CMatrix& mm = this->*memberOf(CGpMTag());
  key.insert(key.end(), mm.getVals(), mm.getVals() + (size_t)mm.getRows() * mm.getCols());
  key.insert(key.end(), pX->getVals(), pX->getVals() + (size_t)pX->getRows() * pX->getCols());
  keyX = pX;
  keyY = py;
}

void CGpB200::upload() const
{
  DIMENSIONMATCH(py->getRows() == pX->getRows());
  int64_t N = pX->getRows();
  int D = (int)pX->getCols(), d = (int)py->getCols();
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
  // The targets are the reference's own m = (y - bias)/scale as CGp::updateM left it (CGp.cpp:248-260).  It is NOT
  // recomputed from the current scale: when the output scale is learnt, CGp::setOptParams writes the new scale and calls
  // updateM() while MupToDate is still true (CGp.cpp:429-437), so m keeps the scale it was last updated with -- the
  // optimiser sees the scale only through the log term and its own gradient slot.  Same here, by construction.
  const CMatrix& mm = this->*memberOf(CGpMTag());
  DIMENSIONMATCH(mm.getRows() == (unsigned int)N && mm.getCols() == (unsigned int)d);
  rc = gpc_set_M(dev, mm.getVals(), N, d, N);
  if(rc)
    fail(rc);
  rc = gpc_ctx_sync(dev); // the caller may change m before the next call
  if(rc)
    fail(rc);
}

void CGpB200::ensureEvaluated() const
{
  if(state == EVALUATED && sameInputs())
    return;
  upload();
  gNat.assign(bridge.getNumParams(), 0.0);
  int rc = gpc_eval(dev, bridge.comps(), bridge.numComps(), 0, evalOut, &gNat[0], 0);
  if(rc)
    fail(rc);
  if(evalOut[2] > 1e-2 && getVerbosity() > 2) // CGp.cpp:883-885
    cout << "Warning: jitter of " << evalOut[2] << " added to K in _updateInvK()." << endl;
  snapshotInputs();
  state = EVALUATED;
  nEvals++;
}

void CGpB200::ensureFactored() const
{
  if(state >= FACTORED && sameInputs())
    return;
  upload();
  int rc = gpc_kern_build(dev, bridge.comps(), bridge.numComps()); // CGp::_updateK (CGp.cpp:693-712)
  if(rc)
    fail(rc);
  double jit = 0.0, logdet = 0.0;
  rc = gpc_jitchol(dev, 20, &jit, &logdet); // LcholK.jitChol(K) (CGp.cpp:882; default maxTries CMatrix.h:1060)
  if(rc)
    fail(rc);
  evalOut[0] = logdet;
  evalOut[2] = jit;
  snapshotInputs();
  state = FACTORED;
}

// ---- sparse approximations (DTC / FITC / DTCVAR) ---------------------------------------------------------------------
bool CGpB200::onDeviceSparse() const
{
  int a = getApproximationType();
  if(!isSparseApproximation() || !(a == DTC || a == FITC || a == DTCVAR))
    return false;
  if(isOptimiseX() || isBackConstrained() || !isSpherical() || isOutputScaleLearnt())
    return false;
  if(!pX || !py || !getKernel() || getNumActive() == 0 || X_u.getRows() != getNumActive())
    return false;
  return bridge.sync(getKernel(), pX->getCols());
}

bool CGpB200::sameSparseInputs() const
{
  if(keyX != pX || keyY != py || !sdev)
    return false;
  const std::vector<double>& p = bridge.naturalParams();
  const CMatrix& mm = this->*memberOf(CGpMTag());
  unsigned int d = getOutputDim();
  size_t nm = (size_t)mm.getRows() * mm.getCols(), nu = (size_t)X_u.getRows() * X_u.getCols();
  size_t nx = (size_t)pX->getRows() * pX->getCols();
  if(skey.size() != p.size() + 2 * d + 1 + nu + nm + nx)
    return false;
  size_t c = 0;
  for(size_t i = 0; i < p.size(); i++)
    if(skey[c++] != p[i])
      return false;
  for(unsigned int j = 0; j < d; j++)
    if(skey[c++] != getScaleVal(j))
      return false;
  for(unsigned int j = 0; j < d; j++)
    if(skey[c++] != getBiasVal(j))
      return false;
  if(skey[c++] != getBetaVal())
    return false;
  const double* uv = X_u.getVals();
  for(size_t i = 0; i < nu; i++)
    if(skey[c++] != uv[i])
      return false;
  const double* mv = mm.getVals();
  for(size_t i = 0; i < nm; i++)
    if(skey[c++] != mv[i])
      return false;
  const double* xv = pX->getVals();
  for(size_t i = 0; i < nx; i++)
    if(skey[c++] != xv[i])
      return false;
  return true;
}

void CGpB200::ensureSparseEvaluated() const
{
  if(sstate == EVALUATED && sameSparseInputs())
    return;
  DIMENSIONMATCH(py->getRows() == pX->getRows());
  int64_t N = pX->getRows();
  int D = (int)pX->getCols(), d = (int)py->getCols(), M = (int)getNumActive(), a = getApproximationType();
  const CMatrix& mm = this->*memberOf(CGpMTag());
  if(!sdev || N != sN || D != sD || d != sd || M != sM || a != sApprox)
  {
    if(sdev)
      gpc_sparse_destroy(sdev);
    sdev = 0;
    // the library takes the reference's enum values (CGp.h:13-19): DTC = 1, FITC = 2, DTCVAR = 4
    int rc = gpc_sparse_create(&sdev, device, a, N, M, D, d);
    if(rc)
      fail(rc);
    sN = N;
    sD = D;
    sd = d;
    sM = M;
    sApprox = a;
    keyX = 0;
  }
  // X and the reference's own m (see upload() for why m is taken as CGp::updateM left it)
  int rc = gpc_sparse_set_data(sdev, pX->getVals(), N, mm.getVals(), N);
  if(rc)
    fail(rc);
  sgNat.assign(bridge.getNumParams(), 0.0);
  sgXu.assign((size_t)M * D, 0.0);
  rc = gpc_sparse_eval(sdev, bridge.comps(), bridge.numComps(), X_u.getVals(), M, getBetaVal(), spOut, &sgNat[0], &sgXu[0],
                       &sgBeta);
  if(rc)
    fail(rc);
  if((spOut[4] > 1e-2 || spOut[5] > 1e-2) && getVerbosity() > 2) // CGp.cpp:778-780, 831-833
    cout << "Warning: jitter of " << (spOut[5] > spOut[4] ? spOut[5] : spOut[4]) << " added to A in updateAD()." << endl;
  skey = bridge.naturalParams();
  for(unsigned int j = 0; j < getOutputDim(); j++)
    skey.push_back(getScaleVal(j));
  for(unsigned int j = 0; j < getOutputDim(); j++)
    skey.push_back(getBiasVal(j));
  skey.push_back(getBetaVal());
  skey.insert(skey.end(), X_u.getVals(), X_u.getVals() + (size_t)X_u.getRows() * X_u.getCols());
  skey.insert(skey.end(), mm.getVals(), mm.getVals() + (size_t)mm.getRows() * mm.getCols());
  skey.insert(skey.end(), pX->getVals(), pX->getVals() + (size_t)pX->getRows() * pX->getCols());
  keyX = pX;
  keyY = py;
  sstate = EVALUATED;
  nEvals++;
}

double CGpB200::logLikelihood() const
{
  if(onDeviceSparse())
  {
    ensureSparseEvaluated();
    // spOut[0] already is -1/2 (...) - d N/2 log 2pi with the reference's constants (CGp.cpp:939-988, 1009-1012)
    return spOut[0] + getKernel()->priorLogProb();
  }
  if(!onDevice())
    return CGp::logLikelihood();
  ensureEvaluated();
  double d = (double)getOutputDim();
  double L = evalOut[1] + d * evalOut[0]; // sum_j m_j' K^-1 m_j + d logdet K (CGp.cpp:920-933)
  if(isOutputScaleLearnt())
    for(unsigned int j = 0; j < getOutputDim(); j++)
      L += 2 * log(fabs(getScaleVal(j))); // CGp.cpp:1002-1008
  L *= -0.5;
  L += getKernel()->priorLogProb();
  L -= d * (double)getNumData() * ndlutil::HALFLOGTWOPI; // CGp.cpp:1012
  return L;
}

double CGpB200::logLikelihoodGradient(CMatrix& g) const
{
  if(onDeviceSparse())
  {
    if(!isMupToDate())
      throw ndlexceptions::Error("updateG() called when M is not updated.");
    ensureSparseEvaluated();
    DIMENSIONMATCH(g.getRows() == 1 && g.getCols() == getOptNumParams());
    g.zeros();
    unsigned int counter = 0;
    if(!isInducingFixed()) // [X_u column-major] (CGp.cpp:1043-1055)
      for(size_t i = 0; i < sgXu.size(); i++)
        g.setVal(sgXu[i], 0, counter++);
    std::vector<double> gk(sgNat);
    bridge.finishGradient(getKernel(), &gk[0]); // transforms and prior gradients, as CKern::getGradTransParams
    for(unsigned int i = 0; i < bridge.getNumParams(); i++)
      g.setVal(gk[i], 0, counter++);
    // log beta: gBeta * gradfact(beta) with the exp transform beta carries (CGp.cpp:1071-1076, CTransform.cpp:50-53)
    g.setVal(sgBeta * getBetaVal(), 0, counter++);
    return logLikelihood();
  }
  if(!onDevice())
    return CGp::logLikelihoodGradient(g);
  if(!isMupToDate()) // CGp::updateG (CGp.cpp:1082-1083)
    throw ndlexceptions::Error("updateG() called when M is not updated.");
  ensureEvaluated();
  unsigned int P = bridge.getNumParams();
  DIMENSIONMATCH(g.getRows() == 1 && g.getCols() == getOptNumParams());
  std::vector<double> gk(gNat);
  bridge.finishGradient(getKernel(), &gk[0]);
  g.zeros();
  unsigned int counter = 0;
  for(unsigned int i = 0; i < P; i++)
    g.setVal(gk[i], 0, counter++);
  if(isOutputScaleLearnt())
  {
    // g_scaleBias = (m' K^-1 m - 1)/scale behind the kernel parameters (CGp.cpp:1062-1069, 1231-1241).  Single output
    // only: with d > 1 the reference's own buffer is 1 x 1 (CGp.cpp:199-202 runs before the flag can be set) and its
    // bound check fails, so onDevice() leaves that case to the inherited code.
    g.setVal(1.0 / getScaleVal(0) * (evalOut[1] - 1.0), 0, counter);
    counter++;
  }
  return logLikelihood();
}

void CGpB200::updateX()
{
  CGp::updateX();
  state = STALE; // *pX changed in place
  sstate = STALE;
}

void CGpB200::posteriorMeanVar(CMatrix& mu, CMatrix& varSigma, const CMatrix& Xin) const
{
  if(onDeviceSparse())
  {
    DIMENSIONMATCH(mu.getCols() == getOutputDim() && varSigma.getCols() == getOutputDim());
    DIMENSIONMATCH(mu.getRows() == Xin.getRows() && varSigma.getRows() == Xin.getRows());
    DIMENSIONMATCH(Xin.getCols() == pX->getCols());
    ensureSparseEvaluated();
    std::vector<double> v(Xin.getRows());
    int rc = gpc_sparse_posterior(sdev, bridge.comps(), bridge.numComps(), Xin.getVals(), Xin.getRows(), Xin.getRows(),
                                  mu.getVals(), &v[0]);
    if(rc)
      fail(rc);
    for(unsigned int j = 0; j < getOutputDim(); j++)
    {
      double scaleVal = getScaleVal(j), biasVal = getBiasVal(j); // CGp.cpp:561-573, 618-626
      for(unsigned int i = 0; i < varSigma.getRows(); i++)
        varSigma.setVal(v[i] * scaleVal * scaleVal, i, j);
      if(scaleVal != 1.0)
        mu.scaleCol(j, scaleVal);
      if(biasVal != 0.0)
        mu.addCol(j, biasVal);
    }
    return;
  }
  if(!onDevice())
  {
    CGp::posteriorMeanVar(mu, varSigma, Xin);
    return;
  }
  DIMENSIONMATCH(mu.getCols() == getOutputDim());
  DIMENSIONMATCH(varSigma.getCols() == getOutputDim());
  DIMENSIONMATCH(mu.getRows() == Xin.getRows());
  DIMENSIONMATCH(varSigma.getRows() == Xin.getRows());
  DIMENSIONMATCH(Xin.getCols() == pX->getCols());
  ensureFactored();
  double quad = 0.0;
  int rc = gpc_solve_alpha(dev, &quad); // CGp::updateAlpha (CGp.cpp:469-484)
  if(rc)
    fail(rc);
  rc = gpc_posterior(dev, bridge.comps(), bridge.numComps(), Xin.getVals(), Xin.getRows(), Xin.getRows(), mu.getVals(),
                     varSigma.getVals());
  if(rc)
    fail(rc);
  for(unsigned int j = 0; j < getOutputDim(); j++)
  {
    for(unsigned int i = 0; i < varSigma.getRows(); i++)
      CHECKZEROORPOSITIVE(varSigma.getVal(i, j) >= 0); // CGp.cpp:607
    double scaleVal = getScaleVal(j), biasVal = getBiasVal(j); // CGp.cpp:561-573, 618-626
    if(scaleVal != 1.0)
    {
      mu.scaleCol(j, scaleVal);
      varSigma.scaleCol(j, scaleVal * scaleVal);
    }
    if(biasVal != 0.0)
      mu.addCol(j, biasVal);
  }
}

void CGpB200::posteriorMean(CMatrix& mu, const CMatrix& Xin) const
{
  if(!onDevice() && !onDeviceSparse())
  {
    CGp::posteriorMean(mu, Xin);
    return;
  }
  CMatrix varSigma(mu.getRows(), mu.getCols());
  posteriorMeanVar(mu, varSigma, Xin);
}

void CGpB200::out(CMatrix& yPred, const CMatrix& Xin) const
{
  DIMENSIONMATCH(yPred.getRows() == Xin.getRows());
  CMatrix muTest(yPred.getRows(), yPred.getCols());
  CMatrix varSigmaTest(yPred.getRows(), yPred.getCols());
  posteriorMeanVar(muTest, varSigmaTest, Xin);
  (this->*memberOf(CGpNoiseTag()))->out(yPred, muTest, varSigmaTest); // pnoise->out (CGp.cpp:451)
}
void CGpB200::out(CMatrix& yPred, CMatrix& probPred, const CMatrix& Xin) const
{
  CMatrix muTest(yPred.getRows(), yPred.getCols());
  CMatrix varSigmaTest(yPred.getRows(), yPred.getCols());
  posteriorMeanVar(muTest, varSigmaTest, Xin);
  (this->*memberOf(CGpNoiseTag()))->out(yPred, probPred, muTest, varSigmaTest); // CGp.cpp:459
}

void CGpB200::download(int which, CMatrix& dst, bool square) const
{
  if(!dev || state == STALE)
    throw ndlexceptions::Error("gpc_b200: nothing evaluated on the device yet");
  unsigned int N = getNumData();
  dst.resize(N, square ? N : getOutputDim());
  int rc = gpc_download(dev, which, dst.getVals(), N);
  if(rc)
    fail(rc);
  if(square && which != GPC_MAT_L)
    dst.setSymmetric(true);
}

CGpB200* readGpB200FromStream(istream& in)
{
  CGpB200* pmodel = new CGpB200();
  pmodel->fromStream(in);
  return pmodel;
}
CGpB200* readGpB200FromFile(const string modelFileName, int verbosity)
{
  if(verbosity > 0)
    cout << "Loading model file." << endl;
  ifstream in(modelFileName.c_str());
  if(!in.is_open())
    throw ndlexceptions::FileReadError(modelFileName);
  CGpB200* pmodel;
  try
  {
    pmodel = readGpB200FromStream(in);
  }
  catch(ndlexceptions::StreamFormatError& err)
  {
    throw ndlexceptions::FileFormatError(modelFileName, err);
  }
  if(verbosity > 0)
    cout << "... done." << endl;
  in.close();
  pmodel->setVerbosity(verbosity);
  return pmodel;
}
