// cgp_b200_check.cpp -- TEST DRIVER (tests/test_gpu_cpp_host.py, tests/test_cpp_host_cpu.py).
//
// Builds the reference's CGp / CGplvm (host LAPACK path, compiled from /root/reference by oracle/build_ref.sh) and the
// drop-in CGpB200 / CGplvmB200 (gpc_b200/cpp, device path through libgpc_b200.so) on the SAME data in one process and
// prints both sets of results as one JSON object: log-likelihood, optimiser-space gradient, predictions through out(),
// and the parameters after a few iterations of the reference's own SCG optimiser driving each class.
//
//   cgp_b200_check gp    N D d seed kern1,kern2,... [scale] [prior] [iters]
//   cgp_b200_check gplvm N q d seed kern1,kern2,... [scale] [prior] [iters]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include "CGpB200.h"
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
    if(prior && c == 0)
    {
      CDist* p = new CGammaDist(); // a prior on the first parameter of the first component
      p->setParam(1.5, 0);
      p->setParam(0.7, 1);
      k->addPrior(p, 0);
    }
    kern.addKern(k);
    delete k;
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
