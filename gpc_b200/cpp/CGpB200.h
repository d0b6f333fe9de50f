// CGpB200.h -- the reference's CGp with its exact-GP (FTC) hot path on the B200.
//
// A CGpB200 IS a CGp (CGp.h:9-482): same constructors, same public members (pX, py, X_u), same optimiser, stream and
// display behaviour -- everything that is not the hot path is inherited, unmodified reference code.  The virtual entry
// points the optimisers and front-ends drive,
//     logLikelihood()               CGp.cpp:913-1013
//     logLikelihoodGradient(g)      CGp.cpp:1016-1144   (updateG, updateCovGradient, CKern::getGradTransParams)
//     out(yPred, X)                 CGp.cpp:445-452     (posteriorMeanVar CGp.cpp:535-663, then the noise model)
// are re-bound to the C ABI of libgpc_b200.so (include/gpc_b200.h): K, LcholK, invK and Alpha live on the device in a
// gpc_ctx owned by this object; one evaluation is ONE gpc_eval call (K build -> jitChol -> K^-1 -> alpha -> gradient,
// one host synchronisation).  The non-virtual methods of the same path (posteriorMeanVar, posteriorMean, the two-output
// out) are redeclared here so that code holding a CGpB200 uses the device too.  CGp::optimise and every optimiser of
// COptimisable (SCG, CG, GD, BFGS) are inherited as they are: they only see the virtuals.
//
// The sparse approximations DTC / FITC / DTCVAR (CGp.cpp:713-861, 939-988, 1146-1413) take the same route through
// gpc_sparse_eval / gpc_sparse_posterior (K_uu, K_uf, the two M x M factorisations, the O(N M^2) products and the gradient
// passes on the device; the optimiser's [X_u][kernel][log beta] layout assembled here).
// Anything outside the device path -- PITC, a learnt output scale on a sparse model, kernel components the library does
// not implement, optimiseX on a CGp -- falls through to the inherited host implementation, call by call.
//
// No reference source is modified:  g++ -include gp_dropin.h gp.cpp  builds the reference's own `gp` front-end on this
// class (build recipe and the resulting gp_l2 / gplvm_l2 executables: INTEGRATION.md, level 2).
#ifndef CGPB200_H
#define CGPB200_H
#include <vector>
#include "CGp.h"
#include "GpcKernBridge.h"
#include "gpc_b200.h"

class CGpB200 : public CGp
{
 public:
  CGpB200();
  CGpB200(CKern* kernel, CNoise* nois, CMatrix* Xin, int approxType = FTC, unsigned int actSetSize = 0, int verbos = 2);
  CGpB200(unsigned int q, unsigned int d, CMatrix* Xin, CMatrix* yin, CKern* kernel, CNoise* nois, int approxType = FTC,
          unsigned int actSetSize = 0, int verbos = 2);
  virtual ~CGpB200();

  // --- virtuals of CProbabilisticOptimisable / CGp / CMapModel
  virtual double logLikelihood() const;
  virtual double logLikelihoodGradient(CMatrix& g) const;
  virtual void updateX();
  virtual void out(CMatrix& yPred, const CMatrix& inData) const;
  // --- non-virtual in CGp: same signatures, device-backed
  void out(CMatrix& yPred, CMatrix& probPred, const CMatrix& inData) const;
  void posteriorMeanVar(CMatrix& mu, CMatrix& varSigma, const CMatrix& X) const;
  void posteriorMean(CMatrix& mu, const CMatrix& X) const;

  // the host copies of the state (for -DDBG style inspection): N x N each, filled from the device on request
  void downloadK(CMatrix& K) const { download(GPC_MAT_K, K, true); }
  void downloadInvK(CMatrix& invK) const { download(GPC_MAT_KINV, invK, true); }
  void downloadLcholK(CMatrix& L) const { download(GPC_MAT_L, L, true); }
  void downloadAlpha(CMatrix& A) const { download(GPC_MAT_ALPHA, A, false); }
  // which device the context is created on (default: $GPC_DEVICE or 0); call before the first evaluation
  void setDevice(int dev);
  // true when the next evaluation will run on the device (FTC, fixed X, every kernel component supported)
  bool onDevice() const;
  // the same for the sparse approximations DTC / FITC / DTCVAR
  bool onDeviceSparse() const;
  // drop the cached evaluation (call after changing *pX or *py in place without going through setOptParams/updateX)
  void invalidate() const { state = STALE; sstate = STALE; }
  // device evaluations so far (one per distinct parameter point)
  unsigned long getNumDeviceEvals() const { return nEvals; }

 private:
  CGpB200(const CGpB200&);            // the object owns a device context: not copyable
  CGpB200& operator=(const CGpB200&);
  enum { STALE = 0, FACTORED = 1, EVALUATED = 2 };
  void init();
  bool sameInputs() const; // kernel parameters, scale, bias, m, data pointers unchanged since the cached evaluation
  void snapshotInputs() const;
  void upload() const;          // *pX and the reference's m = (y - bias)/scale (CGp::updateM, CGp.cpp:248-260) -> device
  void ensureEvaluated() const; // gpc_eval
  void ensureFactored() const;  // K build + jitChol only (prediction does not need K^-1)
  void download(int which, CMatrix& dst, bool square) const;
  void fail(int rc) const;      // rc>0 -> MatrixNonPosDef, rc<0 -> Error(gpc_last_error())
  void ensureSparseEvaluated() const; // gpc_sparse_eval
  bool sameSparseInputs() const;

  mutable gpc_ctx* dev;
  mutable int64_t devN;
  mutable int devD, devd;
  int device;
  mutable GpcKernBridge bridge;
  mutable int state;
  mutable std::vector<double> key; // kernel natural parameters, scales, biases, m at the cached evaluation
  mutable const CMatrix* keyX;
  mutable const CMatrix* keyY;
  mutable double evalOut[3]; // logdet, sum_j m_j' K^-1 m_j, jitter added
  mutable std::vector<double> gNat;
  mutable unsigned long nEvals;
  // sparse approximations
  mutable gpc_sparse* sdev;
  mutable int64_t sN;
  mutable int sM, sD, sd, sApprox;
  mutable int sstate;
  mutable std::vector<double> skey; // kernel parameters, scales, biases, beta, X_u, m
  mutable double spOut[6];
  mutable std::vector<double> sgNat, sgXu;
  mutable double sgBeta;
};

// readGpFromFile (CGp.cpp:1701-1724) for this class
CGpB200* readGpB200FromStream(istream& in);
CGpB200* readGpB200FromFile(const string modelFileName, int verbosity = 2);
#endif
