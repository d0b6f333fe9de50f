// cgp_b200_check.cpp -- TEST DRIVER (tests/test_gpu_cpp_host.py, tests/test_cpp_host_cpu.py).
//
// Builds the reference's CGp / CGplvm (host LAPACK path, compiled from /root/reference by oracle/build_ref.sh) and the
// drop-in CGpB200 / CGplvmB200 (gpc_b200/cpp, device path through libgpc_b200.so) on the SAME data in one process and
// prints both sets of results as one JSON object: log-likelihood, optimiser-space gradient, predictions through out(),
// and the parameters after a few iterations of the reference's own SCG optimiser driving each class.
//
//   cgp_b200_check gp    N D d seed kern1,kern2,... [scale] [prior] [iters]
//   cgp_b200_check gplvm N q d seed kern1,kern2,... [scale] [prior] [iters]
//   cgp_b200_check bridge N D 0 seed kern1,kern2,... 0 [prior]   GpcKernBridge alone (host only)
//   cgp_b200_check download N D d seed kern1,kern2,...                      CGpB200::downloadK/InvK/LcholK/Alpha through identities
//   cgp_b200_check sparse N D d seed kern1,kern2,... approx M beta           the REFERENCE's DTC(1) / FITC(2) / DTCVAR(4) ll + gradient
//   cgp_b200_check sparsedev N D d seed kern1,... approx M beta iters        reference CGp vs CGpB200 on a sparse approximation
//   cgp_b200_check modelwrite N D d seed kern1,kern2,... scale prior path   the REFERENCE writes a gp model file (CGp.cpp:1640-1666)
//   cgp_b200_check modelread 0 0 0 0 path                                   the REFERENCE reads one and prints what it holds
//   cgp_b200_check lvmwrite N q d seed kern1,kern2,... labels 0 path        the REFERENCE writes a gplvm model file (CGplvm.cpp:761-921)
//   cgp_b200_check lvmread 0 0 0 0 path                                     the REFERENCE reads one
//   cgp_b200_check bench N D reps seed kern1,kern2,...      evaluations per second through CGpB200 alone (the metric of
//                                                           bench.py, driven the way COptimisable drives a model)
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <ctime>
#include <string>
#include <vector>
#include "CGpB200.h"
#include "CCmpndKernB200.h"
#include "CGplvmB200.h"

static unsigned long long rngState = 88172645463325252ULL;
static double uniform01()
{
  rngState ^= rngState << 13;
  rngState ^= rngState >> 7;
  rngState ^= rngState << 17;
  return (double)(rngState >> 11) / 9007199254740992.0;
}
static double normal01()
{
  double u = uniform01(), v = uniform01();
  if(u < 1e-300)
    u = 1e-300;
  return sqrt(-2.0 * log(u)) * cos(2.0 * M_PI * v);
}

static CKern* makeComponent(const std::string& name, unsigned int D)
{
  if(name == "rbf") return new CRbfKern(D);
  if(name == "rbfard") return new CRbfardKern(D);
  if(name == "matern32") return new CMatern32Kern(D);
  if(name == "matern52") return new CMatern52Kern(D);
  if(name == "lin") return new CLinKern(D);
  if(name == "poly") return new CPolyKern(D);
  if(name == "white") return new CWhiteKern(D);
  if(name == "bias") return new CBiasKern(D);
  if(name == "ratquad") return new CRatQuadKern(D); // outside the device path: exercises the host fall-through
  fprintf(stderr, "unknown kernel %s\n", name.c_str());
  exit(2);
}

// CComponentKern::components is protected (CKern.h:469-472): member pointer formed inside a derived class
struct DriverComponentPeek : public CComponentKern
{
  static std::vector<CKern*> CComponentKern::*member() { return &DriverComponentPeek::components; }
};

// the same compound kernel twice (one per model), parameters set to a deterministic non-default point
static void buildKernel(CCmpndKern& kern, const std::string& spec, unsigned int D, bool prior)
{
  size_t pos = 0;
  int c = 0;
  while(pos <= spec.size())
  {
    size_t e = spec.find(',', pos);
    if(e == std::string::npos)
      e = spec.size();
    CKern* k = makeComponent(spec.substr(pos, e - pos), D);
    kern.addKern(k);
    if(prior && c == 0)
    {
      // a gamma prior on the first parameter of the first component.  It goes onto the CLONE the compound kernel holds:
      // CKern's copy constructor does not carry priors over, so one added before addKern would be lost
      CKern* held = (static_cast<CComponentKern&>(kern).*DriverComponentPeek::member())[0];
      CDist* p = new CGammaDist();
      p->setParam(1.5, 0);
      p->setParam(0.7, 1);
      held->addPrior(p, 0);
    }
    // k is deliberately NOT deleted: the reference's ARD kernels copy their `scales` matrix with the compiler-generated
    // CMatrix assignment (CKern.cpp:3178), so the clone inside `kern` shares k's buffer (gp.cpp keeps its component
    // objects alive for the whole run for the same reason)
    pos = e + 1;
    c++;
  }
  for(unsigned int i = 0; i < kern.getNumParams(); i++)
  {
    std::string nm = kern.getParamName(i);
    double t = 0.3 * sin(1.0 + 2.0 * i) - 0.2;
    if(nm.find("white") == 0)
      t = -2.5;
    kern.setTransParam(t, i);
  }
}

static void printVec(const char* name, const CMatrix& v, bool last = false)
{
  printf("\"%s\": [", name);
  for(unsigned int i = 0; i < v.getRows() * v.getCols(); i++)
    printf("%s%.17g", i ? ", " : "", v.getVals()[i]);
  printf("]%s\n", last ? "" : ",");
}

static int runGp(unsigned int N, unsigned int D, unsigned int d, const std::string& spec, bool scaleLearnt, bool prior,
                 int iters)
{
  CMatrix X(N, D), y(N, d);
  for(unsigned int j = 0; j < D; j++)
    for(unsigned int i = 0; i < N; i++)
      X.setVal(normal01(), i, j);
  for(unsigned int j = 0; j < d; j++)
    for(unsigned int i = 0; i < N; i++)
      y.setVal(sin(X.getVal(i, 0) + 0.5 * j) + 0.3 * j + 0.1 * normal01(), i, j);
  CMatrix bias(1, d), scale(1, d);
  for(unsigned int j = 0; j < d; j++)
  {
    double s = 0.0;
    for(unsigned int i = 0; i < N; i++)
      s += y.getVal(i, j);
    bias.setVal(s / N, j);
    scale.setVal(scaleLearnt ? 1.3 + 0.2 * j : 1.0, j);
  }
  const unsigned int Ns = 9;
  CMatrix Xs(Ns, D);
  for(unsigned int j = 0; j < D; j++)
    for(unsigned int i = 0; i < Ns; i++)
      Xs.setVal(1.5 * normal01(), i, j);

  CCmpndKern kernRef(X), kernDev(X);
  buildKernel(kernRef, spec, D, prior);
  buildKernel(kernDev, spec, D, prior);
  CGaussianNoise noiseRef(&y), noiseDev(&y);
  CGp ref(&kernRef, &noiseRef, &X, CGp::FTC, 0, 0);
  CGpB200 dev(&kernDev, &noiseDev, &X, CGp::FTC, 0, 0);
  CGp* models[2] = {&ref, &dev};
  for(int k = 0; k < 2; k++)
  {
    models[k]->setBetaVal(1);
    models[k]->setScale(scale);
    models[k]->setBias(bias);
    models[k]->updateM();
    models[k]->setOutputScaleLearnt(scaleLearnt);
    models[k]->setDefaultOptimiser(CGp::SCG);
  }
  printf("{\"mode\": \"gp\", \"N\": %u, \"D\": %u, \"d\": %u, \"on_device\": %d,\n", N, D, d, dev.onDevice() ? 1 : 0);
  CMatrix gRef(1, ref.getOptNumParams()), gDev(1, dev.getOptNumParams());
  // through the base-class pointer: what COptimisable's optimisers call
  double llRef = models[0]->logLikelihoodGradient(gRef);
  double llDev = models[1]->logLikelihoodGradient(gDev);
  printf("\"ll_ref\": %.17g, \"ll_dev\": %.17g, \"ll_dev_again\": %.17g,\n", llRef, llDev, models[1]->logLikelihood());
  printVec("g_ref", gRef);
  printVec("g_dev", gDev);
  unsigned long evalsAfterFirst = dev.getNumDeviceEvals();
  CMatrix yRef(Ns, d), sRef(Ns, d), yDev(Ns, d), sDev(Ns, d), y1Dev(Ns, d);
  ref.out(yRef, sRef, Xs);
  dev.out(yDev, sDev, Xs);
  ((CMapModel*)&dev)->out(y1Dev, Xs); // the virtual one-output form
  printVec("out_ref", yRef);
  printVec("out_dev", yDev);
  printVec("out1_dev", y1Dev);
  printVec("std_ref", sRef);
  printVec("std_dev", sDev);
  if(iters > 0)
  {
    ref.optimise(iters);
    dev.optimise(iters);
    CMatrix pRef(1, ref.getOptNumParams()), pDev(1, dev.getOptNumParams());
    ref.getOptParams(pRef);
    dev.getOptParams(pDev);
    printVec("opt_ref", pRef);
    printVec("opt_dev", pDev);
    printf("\"opt_ll_ref\": %.17g, \"opt_ll_dev\": %.17g,\n", ref.logLikelihood(), dev.logLikelihood());
  }
  printf("\"evals_first\": %lu, \"device_evals\": %lu}\n", evalsAfterFirst, dev.getNumDeviceEvals());
  return 0;
}

static int runGplvm(unsigned int N, unsigned int q, unsigned int d, const std::string& spec, bool scaleLearnt, bool prior,
                    int iters)
{
  // data on a noisy q-dimensional manifold
  CMatrix Y(N, d);
  std::vector<double> Z(N * q), W(q * d);
  for(size_t i = 0; i < Z.size(); i++)
    Z[i] = normal01();
  for(size_t i = 0; i < W.size(); i++)
    W[i] = normal01();
  for(unsigned int j = 0; j < d; j++)
    for(unsigned int i = 0; i < N; i++)
    {
      double v = 0.0;
      for(unsigned int k = 0; k < q; k++)
        v += tanh(Z[i + N * k]) * W[k + q * j];
      Y.setVal(v + 0.5 * j + 0.05 * normal01(), i, j);
    }
  CMatrix Xtmp(1, q);
  CCmpndKern kernRef(Xtmp), kernDev(Xtmp);
  buildKernel(kernRef, spec, q, prior);
  buildKernel(kernDev, spec, q, prior);
  CScaleNoise noiseRef(&Y), noiseDev(&Y);
  CGplvm ref(&kernRef, &noiseRef, q, 0);
  CGplvmB200 dev(&kernDev, &noiseDev, q, 0);
  CGplvm* models[2] = {&ref, &dev};
  for(int k = 0; k < 2; k++)
  {
    models[k]->setInputScaleLearnt(scaleLearnt);
    models[k]->setDefaultOptimiser(CGplvm::SCG);
  }
  printf("{\"mode\": \"gplvm\", \"N\": %u, \"q\": %u, \"d\": %u, \"on_device\": %d,\n", N, q, d, dev.onDevice() ? 1 : 0);
  CMatrix gRef(1, ref.getOptNumParams()), gDev(1, dev.getOptNumParams());
  double llRef = models[0]->logLikelihoodGradient(gRef);
  double llDev = models[1]->logLikelihoodGradient(gDev);
  printf("\"ll_ref\": %.17g, \"ll_dev\": %.17g, \"ll_dev_again\": %.17g,\n", llRef, llDev, models[1]->logLikelihood());
  printVec("g_ref", gRef);
  printVec("g_dev", gDev);
  unsigned long evalsAfterFirst = dev.getNumDeviceEvals();
  const unsigned int Ns = 7;
  CMatrix Xs(Ns, q);
  for(unsigned int j = 0; j < q; j++)
    for(unsigned int i = 0; i < Ns; i++)
      Xs.setVal(0.8 * normal01(), i, j);
  CMatrix muRef(Ns, d), vRef(Ns, d), muDev(Ns, d), vDev(Ns, d);
  ref.posteriorMeanVar(muRef, vRef, Xs);
  dev.posteriorMeanVar(muDev, vDev, Xs);
  printVec("out_ref", muRef);
  printVec("out_dev", muDev);
  printVec("std_ref", vRef);
  printVec("std_dev", vDev);
  if(iters > 0)
  {
    ref.optimise(iters);
    dev.optimise(iters);
    CMatrix pRef(1, ref.getOptNumParams()), pDev(1, dev.getOptNumParams());
    ref.getOptParams(pRef);
    dev.getOptParams(pDev);
    printVec("opt_ref", pRef);
    printVec("opt_dev", pDev);
    printf("\"opt_ll_ref\": %.17g, \"opt_ll_dev\": %.17g,\n", ref.logLikelihood(), dev.logLikelihood());
  }
  printf("\"evals_first\": %lu, \"device_evals\": %lu}\n", evalsAfterFirst, dev.getNumDeviceEvals());
  return 0;
}

// GpcKernBridge on the host alone: the flattened component list, and finishGradient() against the reference's own
// CKern::getGradTransParams (priors + transform factors) for a random symmetric covGrad
static int runBridge(unsigned int N, unsigned int D, const std::string& spec, bool prior)
{
  CMatrix X(N, D);
  for(unsigned int j = 0; j < D; j++)
    for(unsigned int i = 0; i < N; i++)
      X.setVal(normal01(), i, j);
  CCmpndKern kern(X);
  buildKernel(kern, spec, D, prior);
  CMatrix covGrad(N, N);
  for(unsigned int i = 0; i < N; i++)
    for(unsigned int j = 0; j <= i; j++)
    {
      double v = normal01();
      covGrad.setVal(v, i, j);
      covGrad.setVal(v, j, i);
    }
  covGrad.setSymmetric(true);
  CMatrix gTrans(1, kern.getNumParams()), gNat(1, kern.getNumParams());
  kern.getGradTransParams(gTrans, X, covGrad, true); // the reference: priors + gradfact
  kern.getGradParams(gNat, X, covGrad, false);       // natural gradient, no priors: what the device returns
  GpcKernBridge bridge;
  bool ok = bridge.sync(&kern, D);
  printf("{\"mode\": \"bridge\", \"supported\": %d, \"ncomp\": %d, \"nparams\": %u,\n", ok ? 1 : 0, ok ? bridge.numComps() : 0,
         ok ? bridge.getNumParams() : 0);
  if(ok)
  {
    printf("\"types\": [");
    for(int i = 0; i < bridge.numComps(); i++)
      printf("%s%d", i ? ", " : "", bridge.comps()[i].type);
    printf("],\n\"params\": [");
    for(unsigned int i = 0; i < bridge.getNumParams(); i++)
      printf("%s%.17g", i ? ", " : "", bridge.naturalParams()[i]);
    printf("],\n\"kern_params\": [");
    for(unsigned int i = 0; i < kern.getNumParams(); i++)
      printf("%s%.17g", i ? ", " : "", kern.getParam(i));
    printf("],\n");
    std::vector<double> g(gNat.getVals(), gNat.getVals() + kern.getNumParams());
    bridge.finishGradient(&kern, &g[0]);
    CMatrix gBridge(1, kern.getNumParams());
    for(unsigned int i = 0; i < kern.getNumParams(); i++)
      gBridge.setVal(g[i], i);
    printVec("g_bridge", gBridge);
  }
  printVec("g_ref", gTrans, true);
  printf("}\n");
  return 0;
}

// CGpB200::downloadK / downloadInvK / downloadLcholK / downloadAlpha: the device-resident state back in CMatrix form
// (what -DDBG exposes of the reference's K, invK, LcholK, Alpha, CGp.h:359-361), checked through identities
static int runDownload(unsigned int N, unsigned int D, unsigned int d, const std::string& spec)
{
  CMatrix X(N, D), y(N, d);
  for(unsigned int j = 0; j < D; j++)
    for(unsigned int i = 0; i < N; i++)
      X.setVal(normal01(), i, j);
  for(unsigned int j = 0; j < d; j++)
    for(unsigned int i = 0; i < N; i++)
      y.setVal(sin(X.getVal(i, 0) + 0.5 * j) + 0.1 * normal01(), i, j);
  CCmpndKern kern(X);
  buildKernel(kern, spec, D, false);
  CGaussianNoise noise(&y);
  CGpB200 dev(&kern, &noise, &X, CGp::FTC, 0, 0);
  bool threw = false;
  CMatrix K, Kinv, L, A;
  try
  {
    dev.downloadK(K); // nothing evaluated yet: must refuse
  }
  catch(ndlexceptions::Error&)
  {
    threw = true;
  }
  CMatrix g(1, dev.getOptNumParams());
  dev.logLikelihoodGradient(g);
  dev.downloadK(K);
  dev.downloadInvK(Kinv);
  dev.downloadLcholK(L);
  dev.downloadAlpha(A);
  double eK = 0.0, eI = 0.0, eL = 0.0, eA = 0.0, eU = 0.0;
  for(unsigned int i = 0; i < N; i++)
    for(unsigned int j = 0; j < N; j++)
    {
      double kref = (i == j) ? kern.diagComputeElement(X, i) : kern.computeElement(X, i, X, j);
      eK = std::max(eK, fabs(K.getVal(i, j) - kref));
      double s = 0.0, l = 0.0;
      for(unsigned int k = 0; k < N; k++)
      {
        s += K.getVal(i, k) * Kinv.getVal(k, j);
        l += L.getVal(i, k) * L.getVal(j, k);
      }
      eI = std::max(eI, fabs(s - (i == j ? 1.0 : 0.0)));
      eL = std::max(eL, fabs(l - K.getVal(i, j)));
      if(j > i)
        eU = std::max(eU, fabs(L.getVal(i, j))); // LcholK is lower with a zero strict upper triangle (CGp.cpp:890)
    }
  for(unsigned int j = 0; j < d; j++)
    for(unsigned int i = 0; i < N; i++)
    {
      double s = 0.0;
      for(unsigned int k = 0; k < N; k++)
        s += Kinv.getVal(i, k) * (y.getVal(k, j) - dev.getBiasVal(j)) / dev.getScaleVal(j);
      eA = std::max(eA, fabs(s - A.getVal(i, j)));
    }
  printf("{\"mode\": \"download\", \"refused_before_eval\": %d, \"rows\": [%u, %u, %u, %u], \"cols\": [%u, %u, %u, %u], "
         "\"symmetric_flags\": [%d, %d, %d], \"err_K\": %.3g, \"err_KinvK\": %.3g, \"err_LLt\": %.3g, \"err_upper\": %.3g, "
         "\"err_alpha\": %.3g}\n",
         threw ? 1 : 0, K.getRows(), Kinv.getRows(), L.getRows(), A.getRows(), K.getCols(), Kinv.getCols(), L.getCols(),
         A.getCols(), K.isSymmetric() ? 1 : 0, Kinv.isSymmetric() ? 1 : 0, L.isSymmetric() ? 1 : 0, eK, eI, eL, eU, eA);
  return 0;
}

// ---- sparse approximations: the REFERENCE's DTC / DTCVAR / FITC log-likelihood and gradient on seeded inputs, for
// oracle/gp_sparse_oracle.py (CGp.cpp:713-735, 766-861, 939-988, 1244-1413)
static int runSparse(unsigned int N, unsigned int D, unsigned int d, const std::string& spec, int approx, unsigned int M,
                     double betaVal)
{
  CMatrix X(N, D), y(N, d);
  for(unsigned int j = 0; j < D; j++)
    for(unsigned int i = 0; i < N; i++)
      X.setVal(normal01(), i, j);
  for(unsigned int j = 0; j < d; j++)
    for(unsigned int i = 0; i < N; i++)
      y.setVal(sin(X.getVal(i, 0) + 0.5 * j) + 0.1 * normal01(), i, j);
  CCmpndKern kern(X);
  buildKernel(kern, spec, D, false);
  CGaussianNoise noise(&y);
  CGp model(&kern, &noise, &X, approx, M, 0); // picks M rows of X as inducing inputs (CGp::initVals, CGp.cpp:273-284)
  CMatrix bias(1, d);
  for(unsigned int j = 0; j < d; j++)
  {
    double sum = 0.0;
    for(unsigned int i = 0; i < N; i++)
      sum += y.getVal(i, j);
    bias.setVal(sum / N, j);
  }
  model.setBias(bias);
  model.updateM();
  model.setBetaVal(betaVal);
  CMatrix g(1, model.getOptNumParams()), p(1, model.getOptNumParams());
  model.getOptParams(p);
  double ll = model.logLikelihoodGradient(g);
  printf("{\"approx\": \"%s\", \"N\": %u, \"D\": %u, \"d\": %u, \"M\": %u, \"beta\": %.17g, \"ll\": %.17g,\n",
         model.getApproximationStr().c_str(), N, D, d, M, model.getBetaVal(), ll);
  printVec("X", X);
  printVec("y", y);
  printVec("bias", bias);
  printVec("X_u", model.X_u);
  printVec("params", p);
  // prediction through the sparse branches of updateAlpha / _posteriorVar (CGp.cpp:469-534, 600-627)
  const unsigned int Ns = 7;
  CMatrix Xs(Ns, D), mu(Ns, d), var(Ns, d);
  for(unsigned int j = 0; j < D; j++)
    for(unsigned int i = 0; i < Ns; i++)
      Xs.setVal(1.2 * normal01(), i, j);
  model.posteriorMeanVar(mu, var, Xs);
  printVec("Xs", Xs);
  printVec("mu", mu);
  printVec("var", var);
  printVec("g", g, true);
  printf("}\n");
  return 0;
}

// ---- sparse approximations through the drop-in class: reference CGp and CGpB200 with the same approximation, inducing
// inputs and beta on the same seeded data, in one process
static int runSparseDev(unsigned int N, unsigned int D, unsigned int d, const std::string& spec, int approx, unsigned int M,
                        double betaVal, int iters)
{
  CMatrix X(N, D), y(N, d);
  for(unsigned int j = 0; j < D; j++)
    for(unsigned int i = 0; i < N; i++)
      X.setVal(normal01(), i, j);
  for(unsigned int j = 0; j < d; j++)
    for(unsigned int i = 0; i < N; i++)
      y.setVal(sin(X.getVal(i, 0) + 0.5 * j) + 0.1 * normal01(), i, j);
  CMatrix bias(1, d);
  for(unsigned int j = 0; j < d; j++)
  {
    double sum = 0.0;
    for(unsigned int i = 0; i < N; i++)
      sum += y.getVal(i, j);
    bias.setVal(sum / N, j);
  }
  const unsigned int Ns = 7;
  CMatrix Xs(Ns, D);
  for(unsigned int j = 0; j < D; j++)
    for(unsigned int i = 0; i < Ns; i++)
      Xs.setVal(1.2 * normal01(), i, j);
  CCmpndKern kernRef(X), kernDev(X);
  buildKernel(kernRef, spec, D, false);
  buildKernel(kernDev, spec, D, false);
  CGaussianNoise noiseRef(&y), noiseDev(&y);
  CGp ref(&kernRef, &noiseRef, &X, approx, M, 0);
  CGpB200 dev(&kernDev, &noiseDev, &X, approx, M, 0);
  dev.X_u.deepCopy(ref.X_u); // CGp::initVals picks the inducing inputs at random (CGp.cpp:273-284): same ones for both
  CGp* models[2] = {&ref, &dev};
  for(int k = 0; k < 2; k++)
  {
    models[k]->setBias(bias);
    models[k]->updateM();
    models[k]->setBetaVal(betaVal);
    models[k]->setDefaultOptimiser(CGp::SCG);
  }
  printf("{\"mode\": \"sparsedev\", \"approx\": \"%s\", \"N\": %u, \"M\": %u, \"on_device\": %d,\n",
         ref.getApproximationStr().c_str(), N, M, dev.onDeviceSparse() ? 1 : 0);
  CMatrix gRef(1, ref.getOptNumParams()), gDev(1, dev.getOptNumParams());
  double llRef = models[0]->logLikelihoodGradient(gRef);
  double llDev = models[1]->logLikelihoodGradient(gDev);
  printf("\"ll_ref\": %.17g, \"ll_dev\": %.17g, \"ll_dev_again\": %.17g,\n", llRef, llDev, models[1]->logLikelihood());
  printVec("g_ref", gRef);
  printVec("g_dev", gDev);
  unsigned long evalsAfterFirst = dev.getNumDeviceEvals();
  CMatrix muRef(Ns, d), vRef(Ns, d), muDev(Ns, d), vDev(Ns, d);
  ref.posteriorMeanVar(muRef, vRef, Xs);
  dev.posteriorMeanVar(muDev, vDev, Xs);
  printVec("out_ref", muRef);
  printVec("out_dev", muDev);
  printVec("std_ref", vRef);
  printVec("std_dev", vDev);
  if(iters > 0)
  {
    ref.optimise(iters);
    dev.optimise(iters);
    CMatrix pRef(1, ref.getOptNumParams()), pDev(1, dev.getOptNumParams());
    ref.getOptParams(pRef);
    dev.getOptParams(pDev);
    printVec("opt_ref", pRef);
    printVec("opt_dev", pDev);
    printf("\"opt_ll_ref\": %.17g, \"opt_ll_dev\": %.17g,\n", ref.logLikelihood(), dev.logLikelihood());
  }
  printf("\"evals_first\": %lu, \"device_evals\": %lu}\n", evalsAfterFirst, dev.getNumDeviceEvals());
  return 0;
}

// ---- the kernel-class seam: CCmpndKern (host loops, CKern.h:128-157) vs CCmpndKernB200 (gpc_kern_build / gpc_kern_cross)
static int runKern(unsigned int N, unsigned int D, unsigned int N2, const std::string& spec)
{
  CMatrix X(N, D), X2(N2, D);
  for(unsigned int j = 0; j < D; j++)
  {
    for(unsigned int i = 0; i < N; i++)
      X.setVal(normal01(), i, j);
    for(unsigned int i = 0; i < N2; i++)
      X2.setVal(normal01(), i, j);
  }
  CCmpndKern ref(X);
  CCmpndKernB200 dev(X);
  buildKernel(ref, spec, D, false);
  buildKernel(dev, spec, D, false);
  CMatrix Kr(N, N), Kd(N, N), Cr(N, N2), Cd(N, N2);
  const CKern* pr = &ref;
  const CKern* pd = &dev; // through the base class, as CIvm / CGp::posteriorMeanVar call it
  pr->compute(Kr, X);
  pd->compute(Kd, X);
  pr->compute(Cr, X, X2);
  pd->compute(Cd, X, X2);
  CKern* cl = dev.clone();
  CMatrix Kc(N, N);
  cl->compute(Kc, X);
  printf("{\"mode\": \"kern\", \"N\": %u, \"N2\": %u, \"device_builds\": %lu, \"K_maxdiff\": %.17g, \"K2_maxdiff\": %.17g, "
         "\"clone_maxdiff\": %.17g, \"K_symmetric\": %d, \"K_max\": %.17g}\n",
         N, N2, dev.getNumDeviceBuilds(), Kr.maxAbsDiff(Kd), Cr.maxAbsDiff(Cd), Kr.maxAbsDiff(Kc), Kd.isSymmetric() ? 1 : 0,
         Kr.max());
  delete cl;
  return 0;
}

// ---- model files: the reference as the oracle of gpc_gp_model_read / gpc_gp_model_write (tests/test_model_io_cpu.py)
namespace
{
template <typename Tag, typename Tag::type Member> struct DriverPrivateMember
{
  friend typename Tag::type driverMemberOf(Tag) { return Member; }
};
struct DriverNoiseTag
{
  typedef CNoise* CGp::*type;
  friend type driverMemberOf(DriverNoiseTag);
};
template struct DriverPrivateMember<DriverNoiseTag, &CGp::pnoise>;
} // namespace

static void dumpModel(const CGp& model)
{
  printf("{\"num_data\": %u, \"input_dim\": %u, \"output_dim\": %u, \"approx_type\": %d, \"num_active\": %u, "
         "\"learn_scale\": %d, \"learn_bias\": %d,\n",
         model.getNumData(), model.getInputDim(), model.getOutputDim(), model.getApproximationType(), model.getNumActive(),
         model.isOutputScaleLearnt() ? 1 : 0, model.isOutputBiasLearnt() ? 1 : 0);
  printf("\"scale\": [");
  for(unsigned int j = 0; j < model.getOutputDim(); j++)
    printf("%s%.17g", j ? ", " : "", model.getScaleVal(j));
  printf("],\n\"bias\": [");
  for(unsigned int j = 0; j < model.getOutputDim(); j++)
    printf("%s%.17g", j ? ", " : "", model.getBiasVal(j));
  const CKern* kern = model.getKernel();
  printf("],\n\"kern_type\": \"%s\", \"kern_params\": [", kern->getType().c_str());
  for(unsigned int i = 0; i < kern->getNumParams(); i++)
    printf("%s%.17g", i ? ", " : "", kern->getParam(i));
  printf("],\n");
  GpcKernBridge bridge;
  if(bridge.sync(kern, model.getInputDim()))
  {
    printf("\"types\": [");
    for(int i = 0; i < bridge.numComps(); i++)
      printf("%s%d", i ? ", " : "", bridge.comps()[i].type);
    printf("], \"nparams\": [");
    for(int i = 0; i < bridge.numComps(); i++)
      printf("%s%d", i ? ", " : "", bridge.comps()[i].nparams);
    printf("], \"degree\": [");
    for(int i = 0; i < bridge.numComps(); i++)
      printf("%s%.17g", i ? ", " : "", bridge.comps()[i].degree);
    printf("],\n");
  }
  printf("\"prior_log_prob\": %.17g,\n", kern->priorLogProb());
  const CNoise* noise = model.*driverMemberOf(DriverNoiseTag());
  printf("\"noise_type\": \"%s\", \"noise_params\": [", noise->getType().c_str());
  for(unsigned int i = 0; i < noise->getNumParams(); i++)
    printf("%s%.17g", i ? ", " : "", noise->getParam(i));
  printf("]}\n");
}

static int runModelWrite(unsigned int N, unsigned int D, unsigned int d, const std::string& spec, bool scaleLearnt, bool prior,
                         const std::string& path)
{
  CMatrix X(N, D), y(N, d);
  for(unsigned int j = 0; j < D; j++)
    for(unsigned int i = 0; i < N; i++)
      X.setVal(normal01(), i, j);
  for(unsigned int j = 0; j < d; j++)
    for(unsigned int i = 0; i < N; i++)
      y.setVal(sin(X.getVal(i, 0) + 0.5 * j) + 0.1 * normal01(), i, j);
  // "single:rbf" = a model whose kernel is ONE kernel object, not a compound of one (readKernFromStream handles both)
  CCmpndKern cmpnd(X);
  CKern* pk = &cmpnd;
  if(spec.find("single:") == 0)
  {
    pk = makeComponent(spec.substr(7), D);
    for(unsigned int i = 0; i < pk->getNumParams(); i++)
      pk->setTransParam(0.3 * sin(1.0 + 2.0 * i) - 0.2, i);
  }
  else
    buildKernel(cmpnd, spec, D, prior);
  CKern& kern = *pk;
  // values that exercise the text format: an integer, a power of two below one (written 0x1p-2: no '.', the reference
  // reads it back with atoi as 0, CMatrix.cpp:1081-1085), a negative number, a subnormal-free tiny one
  CMatrix bias(1, d), scale(1, d);
  for(unsigned int j = 0; j < d; j++)
  {
    bias.setVal(j == 0 ? -0.23280985888136582 : (j == 1 ? 0.25 : 3.0), j);
    scale.setVal(scaleLearnt ? 1.3 + 0.2 * j : 1.0, j);
  }
  CGaussianNoise noise(&y);
  CGp model(&kern, &noise, &X, CGp::FTC, 0, 0);
  model.setScale(scale);
  model.setBias(bias);
  model.updateM();
  model.setOutputScaleLearnt(scaleLearnt);
  writeGpToFile(model, path, "written by the reference (cgp_b200_check modelwrite)");
  dumpModel(model);
  return 0;
}

static void dumpLvm(const CGplvm& model)
{
  printf("{\"num_data\": %u, \"output_dim\": %d, \"latent_dim\": %d, \"latent_regularised\": %d, \"back_constrained\": %d, "
         "\"dynamics_learnt\": %d, \"has_labels\": %d,\n",
         model.getNumData(), model.getNumProcesses(), model.getLatentDim(), model.isLatentRegularised() ? 1 : 0,
         model.isBackConstrained() ? 1 : 0, model.isDynamicModelLearnt() ? 1 : 0, model.isLabels() ? 1 : 0);
  printf("\"kern_params\": [");
  for(unsigned int i = 0; i < model.pkern->getNumParams(); i++)
    printf("%s%.17g", i ? ", " : "", model.pkern->getParam(i));
  printf("],\n");
  GpcKernBridge bridge;
  if(bridge.sync(model.pkern, model.getLatentDim()))
  {
    printf("\"types\": [");
    for(int i = 0; i < bridge.numComps(); i++)
      printf("%s%d", i ? ", " : "", bridge.comps()[i].type);
    printf("],\n");
  }
  printf("\"noise_type\": \"%s\", \"noise_params\": [", model.pnoise->getType().c_str());
  for(unsigned int i = 0; i < model.pnoise->getNumParams(); i++)
    printf("%s%.17g", i ? ", " : "", model.pnoise->getParam(i));
  printf("],\n\"Y\": [");
  for(int j = 0; j < model.getNumProcesses(); j++)
    for(unsigned int i = 0; i < model.getNumData(); i++)
      printf("%s%.17g", (i || j) ? ", " : "", model.pnoise->getTarget(i, j));
  printf("],\n\"X\": [");
  for(int j = 0; j < model.getLatentDim(); j++)
    for(unsigned int i = 0; i < model.getNumData(); i++)
      printf("%s%.17g", (i || j) ? ", " : "", model.pX->getVal(i, j));
  printf("],\n\"labels\": [");
  if(model.isLabels())
    for(unsigned int i = 0; i < model.getNumData(); i++)
      printf("%s%d", i ? ", " : "", model.getLabel(i));
  printf("]}\n");
}

// the REFERENCE builds a GP-LVM (PCA initialisation), optionally labels it, and writes it with writeGplvmToFile
static int runLvmWrite(unsigned int N, unsigned int q, unsigned int d, const std::string& spec, bool withLabels,
                       const std::string& path)
{
  CMatrix Y(N, d);
  for(unsigned int j = 0; j < d; j++)
    for(unsigned int i = 0; i < N; i++)
      Y.setVal(sin(0.3 * i + j) + (j == 0 && i == 0 ? 0.0 : 0.1 * normal01()) + (i == 1 && j == 1 ? 2.0 : 0.0), i, j);
  Y.setVal(3.0, 2, 0); // an integer and (below) a power of two: written "%a" here, unlike CMatrix::toUnheadedStream
  Y.setVal(0.25, 3, 0);
  CMatrix Xtmp(1, q);
  CCmpndKern kern(Xtmp);
  buildKernel(kern, spec, q, false);
  CScaleNoise noise(&Y);
  CGplvm model(&kern, &noise, q, 0);
  if(withLabels)
  {
    std::vector<int> labels(N);
    for(unsigned int i = 0; i < N; i++)
      labels[i] = (int)(i % 3) - 1;
    model.setLabels(labels);
  }
  writeGplvmToFile(model, path, "written by the reference (cgp_b200_check lvmwrite)");
  dumpLvm(model);
  return 0;
}

static int runLvmRead(const std::string& path)
{
  CGplvm* model = readGplvmFromFile(path, 0);
  dumpLvm(*model);
  return 0;
}

static int runModelRead(const std::string& path)
{
  CGp* model = readGpFromFile(path, 0);
  dumpModel(*model);
  return 0;
}

// setOptParams(theta) -> logLikelihoodGradient(g): one evaluation as SURVEY 8(d) M1 defines it, through the C++ class
static int runBench(unsigned int N, unsigned int D, int reps, const std::string& spec)
{
  CMatrix X(N, D), y(N, 1);
  for(unsigned int j = 0; j < D; j++)
    for(unsigned int i = 0; i < N; i++)
      X.setVal(normal01(), i, j);
  double mean = 0.0;
  for(unsigned int i = 0; i < N; i++)
  {
    y.setVal(sin(X.getVal(i, 0)) + 0.1 * normal01(), i, 0);
    mean += y.getVal(i, 0) / N;
  }
  CCmpndKern kern(X);
  buildKernel(kern, spec, D, false);
  for(unsigned int i = 0; i < kern.getNumParams(); i++) // gamma = 1/D, variance 1, white 0.01: SURVEY 8(d) C2
  {
    std::string nm = kern.getParamName(i);
    double v = 1.0;
    if(nm.find("white") == 0)
      v = 0.01;
    else if(nm.find("inverseWidth") != std::string::npos)
      v = 1.0 / D;
    kern.setParam(v, i);
  }
  CGaussianNoise noise(&y);
  CGpB200 dev(&kern, &noise, &X, CGp::FTC, 0, 0);
  dev.setBiasVal(mean, 0);
  dev.updateM();
  CMatrix theta(1, dev.getOptNumParams()), g(1, dev.getOptNumParams());
  dev.getOptParams(theta);
  double ll = 0.0;
  struct timespec t0, t1;
  for(int r = -3; r < reps; r++)
  {
    if(r == 0)
      clock_gettime(CLOCK_MONOTONIC, &t0);
    theta.setVal(theta.getVal(0) + 1e-9, 0); // a new point every step: nothing is served from the cache
    dev.setOptParams(theta);
    ll = dev.logLikelihoodGradient(g);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  double sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  printf("{\"mode\": \"bench\", \"N\": %u, \"D\": %u, \"reps\": %d, \"ms_per_eval\": %.4f, \"evals_per_sec\": %.3f, "
         "\"ll\": %.12g, \"device_evals\": %lu}\n",
         N, D, reps, 1e3 * sec / reps, reps / sec, ll, dev.getNumDeviceEvals());
  return 0;
}

int main(int argc, char** argv)
{
  if(argc < 7)
  {
    fprintf(stderr, "usage: %s gp|gplvm N D d seed kernels [scale] [prior] [iters]\n", argv[0]);
    return 2;
  }
  std::string mode = argv[1];
  unsigned int N = atoi(argv[2]), D = atoi(argv[3]), d = atoi(argv[4]);
  rngState ^= (unsigned long long)atoll(argv[5]) * 0x9E3779B97F4A7C15ULL;
  for(int i = 0; i < 10; i++)
    uniform01();
  std::string spec = argv[6];
  bool scale = argc > 7 && atoi(argv[7]) != 0;
  bool prior = argc > 8 && atoi(argv[8]) != 0;
  int iters = argc > 9 ? atoi(argv[9]) : 0;
  try
  {
    if(mode == "gp")
      return runGp(N, D, d, spec, scale, prior, iters);
    if(mode == "gplvm")
      return runGplvm(N, D, d, spec, scale, prior, iters);
    if(mode == "download")
      return runDownload(N, D, d, spec);
    if(mode == "sparse") // sparse N D d seed kernels approx(1 dtc, 2 fitc, 4 dtcvar) M beta
      return runSparse(N, D, d, spec, argc > 7 ? atoi(argv[7]) : 1, argc > 8 ? atoi(argv[8]) : 10, argc > 9 ? atof(argv[9]) : 10.0);
    if(mode == "kern") // kern N D N2 seed kernels
      return runKern(N, D, d, spec);
    if(mode == "sparsedev") // sparsedev N D d seed kernels approx M beta iters
      return runSparseDev(N, D, d, spec, argc > 7 ? atoi(argv[7]) : 1, argc > 8 ? atoi(argv[8]) : 10,
                          argc > 9 ? atof(argv[9]) : 10.0, argc > 10 ? atoi(argv[10]) : 0);
    if(mode == "modelwrite")
      return runModelWrite(N, D, d, spec, scale, prior, argc > 9 ? argv[9] : "model.txt");
    if(mode == "modelread")
      return runModelRead(spec);
    if(mode == "lvmwrite")
      return runLvmWrite(N, D, d, spec, scale, argc > 9 ? argv[9] : "lvm.txt");
    if(mode == "lvmread")
      return runLvmRead(spec);
    if(mode == "bridge")
      return runBridge(N, D, spec, prior);
    if(mode == "bench")
      return runBench(N, D, (int)d, spec);
  }
  catch(ndlexceptions::Error& e)
  {
    fprintf(stderr, "ndlexception: %s (%s)\n", e.getMessage().c_str(), e.what());
    return 1;
  }
  catch(std::exception& e)
  {
    fprintf(stderr, "exception: %s\n", e.what());
    return 1;
  }
  fprintf(stderr, "unknown mode %s\n", mode.c_str());
  return 2;
}
