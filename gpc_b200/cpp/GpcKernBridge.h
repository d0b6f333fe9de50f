// GpcKernBridge.h -- host side of the drop-in (C++, compiled against the UNMODIFIED reference headers).
//
// Flattens a reference kernel object (CKern.h: CCmpndKern holding CRbfKern, CRbfardKern, CMatern32Kern, CMatern52Kern,
// CLinKern, CPolyKern, CWhiteKern, CBiasKern) into the gpc_kcomp[] the C ABI takes (include/gpc_b200.h), and finishes a
// natural-parameter gradient the way CKern::getGradTransParams does (CKern.cpp:50-63: priors, then the transform's
// gradient factor).  Only the public interface of the reference classes is used, plus CComponentKern's *protected*
// component list (CKern.h:471) through a derived-class member pointer.
#ifndef GPCKERNBRIDGE_H
#define GPCKERNBRIDGE_H
#include <vector>
#include "CKern.h"
#include "gpc_b200.h"

class GpcKernBridge
{
 public:
  GpcKernBridge() : supported(false), nTotal(0) {}
  // Walks `kern` (component order = parameter order, CKern.h:392-418) and copies the NATURAL parameter values.
  // Returns false when a component is outside the device path (exp, ratquad, mlp, *ard others, tensor, whitefixed):
  // the caller then stays on the reference's host path.
  bool sync(const CKern* kern, unsigned int inputDim);
  bool isSupported() const { return supported; }
  const gpc_kcomp* comps() const { return &kc[0]; }
  int numComps() const { return (int)kc.size(); }
  unsigned int getNumParams() const { return nTotal; }
  const std::vector<double>& naturalParams() const { return vals; }
  // g: 1 x getNumParams() natural-parameter gradient of the log-likelihood (summed over outputs), in/out.
  // Adds each component's prior gradient once (regularise=true on the first output only, CGp.cpp:1105-1112) and
  // multiplies by gradfact of the parameter's transform: the result is what CKern::getGradTransParams returns.
  void finishGradient(const CKern* kern, double* g) const;

 private:
  bool walk(const CKern* kern, unsigned int inputDim);
  bool supported;
  unsigned int nTotal;
  std::vector<const CKern*> parts; // leaf components, parameter order
  std::vector<unsigned int> offs;  // first parameter of each leaf
  std::vector<gpc_kcomp> kc;
  std::vector<double> vals;
};
#endif
