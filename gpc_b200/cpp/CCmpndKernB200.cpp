// CCmpndKernB200.cpp -- see CCmpndKernB200.h.  Builds with the reference's own flags against its unmodified headers.
#include "CCmpndKernB200.h"
#include <cstdlib>

namespace
{
// CComponentKern::components is protected (CKern.h:469-472): member pointer formed inside a derived class
struct ComponentPeek : public CComponentKern
{
  static std::vector<CKern*> CComponentKern::*member() { return &ComponentPeek::components; }
};
} // namespace

void CCmpndKernB200::copyComponents(const CCmpndKern& k)
{
  const std::vector<CKern*>& src = static_cast<const CComponentKern&>(k).*ComponentPeek::member();
  for(size_t i = 0; i < src.size(); i++)
    addKern(src[i]); // clones the component with its parameter values and transforms (CKern.h:382-391)
}

void CCmpndKernB200::init()
{
  dev = 0;
  devN = 0;
  devD = 0;
  const char* e = getenv("GPC_DEVICE");
  device = e ? atoi(e) : 0;
  nBuilds = 0;
}
CCmpndKernB200::~CCmpndKernB200()
{
  if(dev)
    gpc_ctx_destroy(dev);
}
void CCmpndKernB200::setDevice(int d)
{
  if(d != device && dev)
  {
    gpc_ctx_destroy(dev);
    dev = 0;
  }
  device = d;
}

bool CCmpndKernB200::prepare(const CMatrix& X) const
{
  if(getNumKerns() == 0 || !bridge.sync(this, X.getCols()))
    return false;
  int64_t N = X.getRows();
  int D = (int)X.getCols();
  if(!dev || N > devN || D != devD)
  {
    if(dev)
      gpc_ctx_destroy(dev);
    dev = 0;
    if(gpc_ctx_create(&dev, device, N, D, 1) != GPC_OK)
      throw ndlexceptions::Error(std::string("gpc_b200: ") + gpc_last_error());
    devN = N;
    devD = D;
  }
  if(gpc_set_X(dev, X.getVals(), N, D, N) != GPC_OK)
    throw ndlexceptions::Error(std::string("gpc_b200: ") + gpc_last_error());
  return true;
}

void CCmpndKernB200::compute(CMatrix& K, const CMatrix& X) const
{
  DIMENSIONMATCH(K.rowsMatch(X));
  MATRIXPROPERTIES(K.isSquare());
  if(!prepare(X))
  {
    CCmpndKern::compute(K, X);
    return;
  }
  // K(i,i) = diagComputeElement, K(i,j) = computeElement (CKern.h:128-144): gpc_kern_build's semantics
  if(gpc_kern_build(dev, bridge.comps(), bridge.numComps()) != GPC_OK ||
     gpc_download(dev, GPC_MAT_K, K.getVals(), K.getRows()) != GPC_OK)
    throw ndlexceptions::Error(std::string("gpc_b200: ") + gpc_last_error());
  K.setSymmetric(true);
  nBuilds++;
}

void CCmpndKernB200::compute(CMatrix& K, const CMatrix& X, const CMatrix& X2) const
{
  DIMENSIONMATCH(K.rowsMatch(X));
  DIMENSIONMATCH(K.getCols() == X2.getRows());
  DIMENSIONMATCH(X.getCols() == X2.getCols());
  if(!prepare(X))
  {
    CCmpndKern::compute(K, X, X2);
    return;
  }
  // computeElement for every pair (CKern.h:146-157): white noise contributes nothing
  if(gpc_kern_cross(dev, bridge.comps(), bridge.numComps(), X2.getVals(), X2.getRows(), X2.getRows(), K.getVals(),
                    K.getRows()) != GPC_OK)
    throw ndlexceptions::Error(std::string("gpc_b200: ") + gpc_last_error());
  nBuilds++;
}

CKern* readKernB200FromStream(istream& in)
{
  CKern* k = readKernFromStream(in);
  if(k && k->getType() == "cmpnd")
  {
    CCmpndKernB200* b = new CCmpndKernB200(*static_cast<CCmpndKern*>(k));
    delete k;
    return b;
  }
  return k;
}
