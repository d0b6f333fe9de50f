// CGplvmB200.h -- the reference's CGplvm with its hot path on the B200 (twin of CGpB200.h).
//
// A CGplvmB200 IS a CGplvm (CGplvm.h:18-320): constructors, PCA initialisation, optimisers, stream I/O are inherited.
// Re-bound to libgpc_b200.so for the plain GP-LVM (FTC, no dynamics, no back constraints):
//     logLikelihood()              CGplvm.cpp:493-553   (no -dN/2 log 2pi, unlike CGp)
//     logLikelihoodGradient(g)     CGplvm.cpp:555-716   parameter order [kernel][X column-major][scales] (:257-290)
//     posteriorMeanVar(mu,var,X)   CGplvm.cpp:340-362
// One evaluation = one gpc_eval with the dL/dX flag: the N buffers of N x q the reference fills per evaluation
// (CGplvm.cpp:114-115, 569-603) do not exist on the device.  Models with dynamics or back constraints, or with a
// kernel component outside the device path, use the inherited host implementation.
#ifndef CGPLVMB200_H
#define CGPLVMB200_H
#include <vector>
#include "CGplvm.h"
#include "GpcKernBridge.h"
#include "gpc_b200.h"

class CGplvmB200 : public CGplvm
{
 public:
  CGplvmB200();
  CGplvmB200(CKern* kernel, CScaleNoise* nois, const int latDim = 2, const int verbos = 2);
  CGplvmB200(CKern* kernel, CKern* dynKernel, CScaleNoise* nois, const int latDim = 2, const int verbos = 2);
  CGplvmB200(CKern* kernel, CMatrix* backKernel, CScaleNoise* nois, const int latDim = 2, const int verbos = 2);
  CGplvmB200(CKern* kernel, CKern* dynKernel, CMatrix* backKernel, CScaleNoise* nois, const int latDim = 2,
             const int verbos = 2);
  virtual ~CGplvmB200();

  virtual double logLikelihood() const;
  virtual double logLikelihoodGradient(CMatrix& g) const;
  virtual void setOptParams(const CMatrix& param);
  virtual void updateX();
  void posteriorMeanVar(CMatrix& mu, CMatrix& varSigma, const CMatrix& X) const;

  void setDevice(int dev);
  bool onDevice() const;
  void invalidate() const { fresh = false; }
  unsigned long getNumDeviceEvals() const { return nEvals; }

 private:
  CGplvmB200(const CGplvmB200&);            // the object owns a device context: not copyable
  CGplvmB200& operator=(const CGplvmB200&);
  void init();
  void ensureEvaluated() const;
  void fail(int rc) const;
  mutable gpc_ctx* dev;
  mutable int64_t devN;
  mutable int devD, devd;
  int device;
  mutable GpcKernBridge bridge;
  mutable bool fresh;
  mutable std::vector<double> key; // kernel natural parameters of the cached evaluation
  mutable double evalOut[3];
  mutable std::vector<double> gNat, gLatent;
  mutable unsigned long nEvals;
};

CGplvmB200* readGplvmB200FromStream(istream& in);
CGplvmB200* readGplvmB200FromFile(const string modelFileName, const int verbosity = 2);
#endif
