// GpcKernBridge.cpp -- see GpcKernBridge.h.  Builds with the reference's own flags (-std=gnu++98).
#include "GpcKernBridge.h"

namespace
{
// CComponentKern::components is protected (CKern.h:469-472).  A member pointer formed inside a derived class is the
// standard way to read it from an object we are handed as a plain CComponentKern.
struct ComponentPeek : public CComponentKern
{
  static std::vector<CKern*> CComponentKern::*member() { return &ComponentPeek::components; }
};

int typeOf(const std::string& t)
{
  if(t == "white") return GPC_KERN_WHITE;       // CKern.cpp:635
  if(t == "bias") return GPC_KERN_BIAS;         // CKern.cpp:921
  if(t == "rbf") return GPC_KERN_RBF;           // CKern.cpp:1060
  if(t == "rbfard") return GPC_KERN_RBFARD;     // CKern.cpp:3191
  if(t == "matern32") return GPC_KERN_MATERN32; // CKern.cpp:1740
  if(t == "matern52") return GPC_KERN_MATERN52; // CKern.cpp:1987
  if(t == "lin") return GPC_KERN_LIN;           // CKern.cpp:2251
  if(t == "poly") return GPC_KERN_POLY;         // CKern.cpp:2713
  return -1;
}
} // namespace

bool GpcKernBridge::walk(const CKern* kern, unsigned int inputDim)
{
  const std::string t = kern->getType();
  if(t == "cmpnd") // CCmpndKern: the sum of its components (CKern.cpp:219-226); a sum of sums is a sum
  {
    const CComponentKern* ck = dynamic_cast<const CComponentKern*>(kern);
    if(!ck)
      return false;
    const std::vector<CKern*>& list = ck->*ComponentPeek::member();
    for(size_t i = 0; i < list.size(); i++)
      if(!walk(list[i], inputDim))
        return false;
    return true;
  }
  int type = typeOf(t);
  if(type < 0 || (int)kc.size() >= GPC_MAX_COMPONENTS)
    return false;
  if(kern->getNumParams() != (unsigned int)gpc_kern_nparams(type, (int)inputDim))
    return false;
  gpc_kcomp c;
  c.type = type;
  c.nparams = (int)kern->getNumParams();
  c.params = 0;
  c.degree = 2.0;
  if(type == GPC_KERN_POLY)
  {
    const CPolyKern* pk = dynamic_cast<const CPolyKern*>(kern);
    if(!pk)
      return false;
    c.degree = pk->getDegree(); // fixed, not a parameter (CKern.cpp:2723-2729)
  }
  parts.push_back(kern);
  offs.push_back(nTotal);
  for(unsigned int i = 0; i < kern->getNumParams(); i++)
    vals.push_back(kern->getParam(i));
  nTotal += kern->getNumParams();
  kc.push_back(c);
  return true;
}

bool GpcKernBridge::sync(const CKern* kern, unsigned int inputDim)
{
  parts.clear();
  offs.clear();
  kc.clear();
  vals.clear();
  nTotal = 0;
  supported = kern && walk(kern, inputDim) && !kc.empty() && nTotal == kern->getNumParams() && nTotal <= GPC_MAX_PARAMS;
  if(supported)
    for(size_t i = 0; i < kc.size(); i++)
      kc[i].params = &vals[offs[i]];
  return supported;
}

void GpcKernBridge::finishGradient(const CKern* kern, double* g) const
{
  // priors live in the components (CCmpndKern refuses them, CKern.h:439-442); each leaf adds its own at the end of
  // getGradParams when regularise is set (e.g. CKern.cpp:1238-1239)
  for(size_t i = 0; i < parts.size(); i++)
  {
    if(parts[i]->getNumPriors() == 0)
      continue;
    CMatrix sub(1, parts[i]->getNumParams());
    for(unsigned int p = 0; p < parts[i]->getNumParams(); p++)
      sub.setVal(g[offs[i] + p], p);
    parts[i]->addPriorGrad(sub);
    for(unsigned int p = 0; p < parts[i]->getNumParams(); p++)
      g[offs[i] + p] = sub.getVal(p);
  }
  // natural -> transformed parameters (CKern.cpp:55-62)
  for(unsigned int i = 0; i < kern->getNumTransforms(); i++)
  {
    unsigned int idx = kern->getTransformIndex(i);
    g[idx] *= kern->getTransformGradFact(kern->getParam(idx), i);
  }
}
