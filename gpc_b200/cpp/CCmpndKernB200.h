// CCmpndKernB200.h -- the reference's compound kernel with its two matrix-valued `compute` virtuals on the B200.
//
// CKern::compute(K, X) and compute(K, X, X2) (CKern.h:128-157) are the O(N^2) double-virtual-call loops behind every
// kernel matrix the reference builds outside CGp::_updateK: prediction (CGp.cpp:540-545, also for the sparse models),
// CIvm::updateK / posterior (CIvm.cpp:131), CGplvm::posteriorMeanVar (CGplvm.cpp:348), the back-constraint kernel of
// gplvm.cpp:530.  A CCmpndKernB200 IS a CCmpndKern (same constructors, addKern, parameters, transforms, priors, stream
// format); only those two virtuals are re-bound: the components are flattened by GpcKernBridge and the matrix is built by
// gpc_kern_build / gpc_kern_cross (kbuild_kernel / kcross_kernel), then copied into the caller's CMatrix.  A compound
// holding a component outside the device path falls through to the inherited loops.
//
//     g++ ... -include ivm_dropin.h -c ivm.cpp      builds the reference's IVM front-end on this class (no source change)
#ifndef CCMPNDKERNB200_H
#define CCMPNDKERNB200_H
#include "CKern.h"
#include "GpcKernBridge.h"
#include "gpc_b200.h"

class CCmpndKernB200 : public CCmpndKern
{
 public:
  CCmpndKernB200() : CCmpndKern() { init(); }
  CCmpndKernB200(unsigned int inDim) : CCmpndKern(inDim) { init(); }
  CCmpndKernB200(const CMatrix& X) : CCmpndKern(X) { init(); }
  // Copies are built component by component through addKern (which clones).  The reference's own copy constructor cannot be
  // used: CCmpndKern::CCmpndKern(const CCmpndKern&) (CKern.cpp:142-148) copies the component vector and then appends a
  // clone of every element while iterating over the growing vector -- it never terminates (nothing in the reference
  // copies a compound kernel, so it went unnoticed).
  CCmpndKernB200(const CCmpndKern& k) : CCmpndKern(k.getInputDim()) { init(); copyComponents(k); }
  CCmpndKernB200(const CCmpndKernB200& k) : CCmpndKern(k.getInputDim()) { init(); copyComponents(k); }
  virtual ~CCmpndKernB200();
  CCmpndKernB200* clone() const { return new CCmpndKernB200(*this); }

  virtual void compute(CMatrix& K, const CMatrix& X) const;
  virtual void compute(CMatrix& K, const CMatrix& X, const CMatrix& X2) const;
  using CCmpndKern::compute; // the per-row / index forms stay the reference's

  void setDevice(int d);
  unsigned long getNumDeviceBuilds() const { return nBuilds; }

 private:
  CCmpndKernB200& operator=(const CCmpndKernB200&);
  void init();
  void copyComponents(const CCmpndKern& k);
  bool prepare(const CMatrix& X) const; // false: stay on the host path
  mutable gpc_ctx* dev;
  mutable int64_t devN;
  mutable int devD;
  int device;
  mutable GpcKernBridge bridge;
  mutable unsigned long nBuilds;
};

// readKernFromStream (CKern.cpp:4192-4259) for callers that read a kernel themselves: a compound kernel comes back as a
// CCmpndKernB200 (same components and parameters), anything else as the reference created it
CKern* readKernB200FromStream(istream& in);
#endif
