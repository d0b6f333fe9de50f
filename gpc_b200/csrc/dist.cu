// dist.cu -- the multi-GPU exact-GP evaluation: K, its Cholesky factor and K^-1 are never held by one
// GPU.  The N x N matrix is cut into nb x nb blocks dealt 2-D block-cyclically to a P x Q process grid (rank = p Q + q
// owns global block (i, j) iff i % P == p and j % Q == q); every rank stores only its blocks on or below the diagonal
// as ONE local ML x NL matrix, and K -> K^-1 happens IN PLACE in a single right-looking sweep that fuses the three
// LAPACK stages the reference calls one after the other (dpotrf_, then dpotri_ = dtrtri + dlauum; CMatrix.cpp:371-432,
// lapack.h:59-73):
//
//   at step k the local matrix holds   [ Kinv~ (i,j <= k) |                ]   Kinv~ : partial sums of W'W
//                                      [ B     (i > k, j <= k) | A (i,j > k) ]   B : rows of L^-1 under construction
//                                                                              A : the Schur complement (as dpotrf_)
//   panel k:   L_kk, W_kk = L_kk^-1          (owner of block (k,k): potrf_inv_rec, the single-GPU kernels)
//              L_ik = A_ik W_kk'  (i > k)    (owners of block column k: one GEMM each)
//              R_kj = W_kk B_kj   (j < k)    (owners of block row k: one GEMM each)
//   the panel is ONE stack of blocks  S_g = L_gk (g > k),  W_kk' (g = k),  R_kg' (g < k)   and every block of the matrix
//   gets the same rank-nb update      T_ij <- beta_ij T_ij - sgn_i S_i S_j'      sgn_i = +1 (i > k), -1 (i <= k),
//                                                                                beta = 0 in block row / column k
//   which is the Cholesky trailing update (i,j > k), the forward substitution L W = I (i > k, j <= k) and the W'W
//   accumulation (i,j <= k) at once.  Every step costs N^2 nb flops whatever k is, so ANY cyclic ownership is load
//   balanced, and after the last step the local matrix IS the local part of K^-1 (N^3 flops in total, as the reference).
//
// The panel is produced already in the form the tensor-core engine consumes (ozaki.cu: S int8 planes + row scales per
// block = a "slot"), so what travels between GPUs is the sliced panel (8 N nb bytes + scales per step, the same as the
// fp64 panel) and no consumer slices anything.  Collectives per step: one broadcast of W_kk (nb x nb), one grouped
// broadcast of the N / nb slots (each from the rank that produced it).  Look-ahead of one panel: block row / column k+1
// is updated first on a high-priority stream and panel k+1 is produced and sent while the bulk update of step k still runs.
// log det, alpha = K^-1 m (symmetric block product + all-reduce) and the gradient partial sums (fused pass over the local
// blocks + all-reduce of P doubles) follow.  Two communication back-ends behind one interface:
//   NCCL   one process per GPU (torchrun / MPI-style launch): ncclBroadcast / ncclAllReduce on the rank's own
//          communicator (libnccl.so.2 is loaded at run time, the library has no link-time dependency on it);
//   local  ONE process driving ndev devices with a worker thread per device: peer copies (cudaMemcpyPeerAsync over
//          NVLink) ordered by CUDA events -- what SURVEY 8(b)'s gpc_ctx_create(devices, ndev, ...) asks for.  The same
//          device may be listed several times (tests of the block-cyclic logic on a single GPU).
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include "common.cuh"

#define GPC_CHECK(expr)            \
  do {                             \
    int _rc = (expr);              \
    if (_rc != GPC_OK) return _rc; \
  } while (0)

namespace gpc {

// ------------------------------------------------------------------------------------------------------
// communication back-ends
// ------------------------------------------------------------------------------------------------------
struct BcastItem {
  void* buf;
  size_t bytes;
  int root;
};

struct Comm {
  int rank = 0, world = 1;
  virtual ~Comm() {}
  // every rank passes the SAME list (its own addresses); on return the copies are queued on stream s
  virtual int bcast_group(const BcastItem* items, int n, cudaStream_t s) = 0;
  // in-place all-reduce of n doubles in device memory (sum or max); result valid in stream order on s
  virtual int allreduce(double* dev, int n, bool is_max, cudaStream_t s) = 0;
  virtual const char* name() const = 0;
};

struct SoloComm : Comm {
  int bcast_group(const BcastItem*, int, cudaStream_t) override { return GPC_OK; }
  int allreduce(double*, int, bool, cudaStream_t) override { return GPC_OK; }
  const char* name() const override { return "single"; }
};

// ---- NCCL through dlopen ------------------------------------------------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (api.handle) return &api;
  const char* names[] = {getenv("GPC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error(std::string("NCCL back-end: cannot load libnccl.so.2 (") + (dlerror() ? dlerror() : "?") + ")");
    return nullptr;
  }
#define GPC_NCCL_SYM(field, sym)                                   \
  *(void**)(&api.field) = dlsym(h, sym);                           \
  if (!api.field) {                                                \
    set_error(std::string("NCCL back-end: missing symbol ") + sym); \
    return nullptr;                                                \
  }
  GPC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  GPC_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  GPC_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  GPC_NCCL_SYM(GroupStart, "ncclGroupStart")
  GPC_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  GPC_NCCL_SYM(Broadcast, "ncclBroadcast")
  GPC_NCCL_SYM(AllReduce, "ncclAllReduce")
  GPC_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  GPC_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef GPC_NCCL_SYM
  api.handle = h;
  return &api;
}
#define GPC_NCCL_CHECK(expr)                                                                              \
  do {                                                                                                    \
    ncclResult_t _r = (expr);                                                                             \
    if (_r != ncclSuccess) {                                                                              \
      set_error(std::string(#expr) + ": " + api->GetErrorString(_r) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
      return GPC_ERR_CUDA;                                                                                \
    }                                                                                                     \
  } while (0)

struct NcclComm : Comm {
  NcclApi* api = nullptr;
  ncclComm_t comm = nullptr;
  ~NcclComm() override {
    if (comm && api) api->CommDestroy(comm);
  }
  int bcast_group(const BcastItem* items, int n, cudaStream_t s) override {
    if (n <= 0) return GPC_OK;
    GPC_NCCL_CHECK(api->GroupStart());
    for (int i = 0; i < n; i++)
      GPC_NCCL_CHECK(api->Broadcast(items[i].buf, items[i].buf, items[i].bytes, ncclChar, items[i].root, comm, s));
    GPC_NCCL_CHECK(api->GroupEnd());
    return GPC_OK;
  }
  int allreduce(double* dev, int n, bool is_max, cudaStream_t s) override {
    GPC_NCCL_CHECK(api->AllReduce(dev, dev, (size_t)n, ncclDouble, is_max ? ncclMax : ncclSum, comm, s));
    return GPC_OK;
  }
  const char* name() const override { return "nccl"; }
};

// ---- one process, one worker thread per device: peer copies ordered by events ----------------------------------
struct LocalHub {
  int world = 0;
  std::vector<int> dev;
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  long generation = 0;
  bool failed = false;  // a rank left the protocol with an error: nobody may block on it any more
  std::vector<cudaEvent_t> e_ready, e_done;
  std::vector<std::vector<void*>> ptrs;  // [rank][item]
  std::vector<double*> host;             // pinned staging of the all-reduce, one per rank
  size_t host_cap = 0;
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (failed) return;
    long gen = generation;
    if (++waiting == world) {
      waiting = 0;
      generation++;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != gen || failed; });
    }
  }
  void fail() {
    std::lock_guard<std::mutex> lk(mu);
    failed = true;
    cv.notify_all();
  }
};

struct LocalComm : Comm {
  LocalHub* hub = nullptr;
  int bcast_group(const BcastItem* items, int n, cudaStream_t s) override {
    if (n <= 0) return GPC_OK;
    LocalHub& h = *hub;
    GPC_CUDA_CHECK(cudaEventRecord(h.e_ready[rank], s));
    h.ptrs[rank].resize((size_t)n);
    for (int i = 0; i < n; i++) h.ptrs[rank][(size_t)i] = items[i].buf;
    h.barrier();
    if (h.failed) return GPC_ERR_STATE;
    bool am_root = false;
    for (int i = 0; i < n; i++) {
      const int root = items[i].root;
      if (root == rank) {
        am_root = true;
        continue;
      }
      GPC_CUDA_CHECK(cudaStreamWaitEvent(s, h.e_ready[root], 0));
      GPC_CUDA_CHECK(cudaMemcpyPeerAsync(items[i].buf, h.dev[rank], h.ptrs[root][(size_t)i], h.dev[root], items[i].bytes, s));
    }
    GPC_CUDA_CHECK(cudaEventRecord(h.e_done[rank], s));
    h.barrier();
    if (h.failed) return GPC_ERR_STATE;
    if (am_root)  // the sources may only be overwritten once every reader has taken its copy
      for (int r = 0; r < world; r++)
        if (r != rank) GPC_CUDA_CHECK(cudaStreamWaitEvent(s, h.e_done[r], 0));
    h.barrier();  // all waits are queued before anybody re-records the events
    return h.failed ? GPC_ERR_STATE : GPC_OK;
  }
  int allreduce(double* dev, int n, bool is_max, cudaStream_t s) override {
    LocalHub& h = *hub;
    if ((size_t)n > h.host_cap) {
      set_error("local all-reduce: staging buffer too small");
      return GPC_ERR_ARG;
    }
    GPC_CUDA_CHECK(cudaMemcpyAsync(h.host[rank], dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    GPC_CUDA_CHECK(cudaStreamSynchronize(s));
    h.barrier();
    if (h.failed) return GPC_ERR_STATE;
    double* out = h.host[rank] + h.host_cap;  // second half of this rank's staging: the reduced vector
    for (int i = 0; i < n; i++) {
      double v = h.host[0][i];
      for (int r = 1; r < world; r++) v = is_max ? (h.host[r][i] > v ? h.host[r][i] : v) : v + h.host[r][i];
      out[i] = v;
    }
    h.barrier();  // everybody has read every input
    GPC_CUDA_CHECK(cudaMemcpyAsync(dev, out, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    return h.failed ? GPC_ERR_STATE : GPC_OK;
  }
  const char* name() const override { return "local"; }
};

// ------------------------------------------------------------------------------------------------------
// per-rank state and the algorithm
// ------------------------------------------------------------------------------------------------------
static const int DS_LOGDET = 0, DS_QUAD = 1, DS_INFO = 2, DS_G = 8;  // layout of the device / host scalar vectors

struct DistRank {
  int device = 0;
  Comm* comm = nullptr;
  int P = 1, Q = 1, p = 0, q = 0, rank = 0, world = 1;
  int64_t N = 0, Np = 0;
  int nb = 0, NBt = 0, D = 0, d = 0, S = 8;
  int nlr = 0, nlc = 0;      // local block rows / columns
  int64_t ML = 0, NL = 0;    // local matrix
  double *X = nullptr, *M = nullptr, *alpha = nullptr, *y = nullptr, *T = nullptr;
  double *Wb = nullptr, *Lp = nullptr, *tmpL = nullptr, *Tpool = nullptr;
  double *Wb2 = nullptr, *LF[2] = {nullptr, nullptr};  // single-rank fast chain: second W_kk buffer, the fp64 tiles L_{k+1,k}
  cudaStream_t s_chain = nullptr;
  SmPartition* part = nullptr;  // single rank: the chain stream owns a few SMs, the panel / main streams the rest
  int* emax = nullptr;
  uint8_t* slots[2] = {nullptr, nullptr};
  OzCycMaps maps[2];
  double* scal = nullptr;     // device scalars
  int* info = nullptr;        // device
  int* errflag = nullptr;     // device (tensor-core kernel protocol errors)
  double* partial = nullptr;  // gradient partial sums
  double* hres = nullptr;     // pinned
  int max_ctas = 0;
  cudaStream_t s_main = nullptr, s_panel = nullptr, s_comm = nullptr;
  std::vector<cudaEvent_t> ev;  // 5 per step
  cudaEvent_t tev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int64_t launches = 0;
  double last_ms[5] = {0, 0, 0, 0, 0};
  bool haveData = false, haveInv = false;
  int result = GPC_OK;
  std::string err;

  CycMap cyc() const {
    CycMap cm;
    cm.on = 1;
    cm.nb = nb;
    cm.P = P;
    cm.Q = Q;
    cm.p = p;
    cm.q = q;
    cm.ML = ML;
    cm.NL = NL;
    return cm;
  }
  int owner(int i, int j) const { return (i % P) * Q + (j % Q); }
  // the rank that produces block g of panel k
  int producer(int g, int k) const { return g > k ? owner(g, k) : (g < k ? owner(k, g) : owner(k, k)); }
  // number of local block rows with global index < g (first local row block with global index >= g)
  int lrow_lb(int g) const { return g <= p ? 0 : (g - p + P - 1) / P; }
  int lcol_lb(int g) const { return g <= q ? 0 : (g - q + Q - 1) / Q; }
};

static int rank_alloc(DistRank& r) {
  GPC_CUDA_CHECK(cudaSetDevice(r.device));
  cudaDeviceProp prop;
  GPC_CUDA_CHECK(cudaGetDeviceProperties(&prop, r.device));
  if (prop.major < 10) {
    set_error("gpc_b200 is built for sm_100a only");
    return GPC_ERR_CUDA;
  }
  r.max_ctas = prop.multiProcessorCount * 2;
  const size_t np = (size_t)r.Np, nb = (size_t)r.nb;
  GPC_CUDA_CHECK(cudaMalloc(&r.X, np * r.D * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&r.M, np * r.d * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&r.alpha, np * r.d * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&r.y, np * r.d * sizeof(double)));
  if (r.ML > 0 && r.NL > 0) GPC_CUDA_CHECK(cudaMalloc(&r.T, (size_t)r.ML * (size_t)r.NL * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&r.Wb, nb * nb * sizeof(double)));
  GPC_CUDA_CHECK(cudaMemset(r.Wb, 0, nb * nb * sizeof(double)));
  if (r.world == 1) {
    GPC_CUDA_CHECK(cudaMalloc(&r.Wb2, nb * nb * sizeof(double)));
    GPC_CUDA_CHECK(cudaMemset(r.Wb2, 0, nb * nb * sizeof(double)));
    for (int b = 0; b < 2; b++) GPC_CUDA_CHECK(cudaMalloc(&r.LF[b], nb * nb * sizeof(double)));
  }
  const size_t strip = (size_t)(r.ML > r.NL ? r.ML : r.NL) + nb;
  GPC_CUDA_CHECK(cudaMalloc(&r.Lp, strip * nb * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&r.emax, strip * sizeof(int)));
  GPC_CUDA_CHECK(cudaMalloc(&r.tmpL, (nb / 2 + TILE) * (nb / 2 + TILE) * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&r.Tpool, (potrf_inv_tspace((int64_t)nb) + 16) * sizeof(double)));
  const size_t sb = oz_slot_bytes(r.nb, r.S);
  for (int b = 0; b < 2; b++) {
    GPC_CUDA_CHECK(cudaMalloc(&r.slots[b], sb * (size_t)r.NBt));
    GPC_CUDA_CHECK(cudaMemset(r.slots[b], 0, sb * (size_t)r.NBt));
    GPC_CHECK(oz_cyc_maps(&r.maps[b], r.slots[b], r.NBt, r.nb, r.S));
  }
  GPC_CUDA_CHECK(cudaMalloc(&r.scal, (DS_G + GPC_MAX_PARAMS) * sizeof(double)));
  GPC_CUDA_CHECK(cudaMalloc(&r.info, sizeof(int)));
  GPC_CUDA_CHECK(cudaMalloc(&r.errflag, sizeof(int)));
  GPC_CUDA_CHECK(cudaMemset(r.errflag, 0, sizeof(int)));
  GPC_CUDA_CHECK(cudaMalloc(&r.partial, (size_t)r.max_ctas * GPC_MAX_PARAMS * sizeof(double)));
  GPC_CUDA_CHECK(cudaMallocHost(&r.hres, (DS_G + GPC_MAX_PARAMS) * sizeof(double)));
  int lo = 0, hi = 0;
  GPC_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = numerically lowest = highest priority
  if (r.world == 1) r.part = smpart_create(r.device, 0, 0);
  if (r.part) {
    GPC_CHECK(smpart_stream(r.part, false, lo, &r.s_main));
    GPC_CHECK(smpart_stream(r.part, false, hi, &r.s_panel));
    GPC_CHECK(smpart_stream(r.part, true, hi, &r.s_chain));
  } else {
    GPC_CUDA_CHECK(cudaStreamCreateWithPriority(&r.s_main, cudaStreamNonBlocking, lo));
    GPC_CUDA_CHECK(cudaStreamCreateWithPriority(&r.s_panel, cudaStreamNonBlocking, hi));
    GPC_CUDA_CHECK(cudaStreamCreateWithPriority(&r.s_chain, cudaStreamNonBlocking, hi));
  }
  GPC_CUDA_CHECK(cudaStreamCreateWithPriority(&r.s_comm, cudaStreamNonBlocking, hi));
  r.ev.resize((size_t)8 * r.NBt + 8);
  for (auto& e : r.ev) GPC_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (int i = 0; i < 6; i++) GPC_CUDA_CHECK(cudaEventCreate(&r.tev[i]));
  return GPC_OK;
}

static void rank_free(DistRank& r) {
  cudaSetDevice(r.device);
  if (r.s_main) cudaStreamSynchronize(r.s_main);
  if (r.s_panel) cudaStreamSynchronize(r.s_panel);
  if (r.s_comm) cudaStreamSynchronize(r.s_comm);
  if (r.s_chain) cudaStreamSynchronize(r.s_chain);
  cudaFree(r.Wb2); cudaFree(r.LF[0]); cudaFree(r.LF[1]);
  cudaFree(r.X); cudaFree(r.M); cudaFree(r.alpha); cudaFree(r.y); cudaFree(r.T); cudaFree(r.Wb); cudaFree(r.Lp);
  cudaFree(r.emax); cudaFree(r.tmpL); cudaFree(r.Tpool); cudaFree(r.slots[0]); cudaFree(r.slots[1]); cudaFree(r.scal);
  cudaFree(r.info); cudaFree(r.errflag); cudaFree(r.partial);
  if (r.hres) cudaFreeHost(r.hres);
  for (auto e : r.ev)
    if (e) cudaEventDestroy(e);
  for (int i = 0; i < 6; i++)
    if (r.tev[i]) cudaEventDestroy(r.tev[i]);
  if (r.s_main) cudaStreamDestroy(r.s_main);
  if (r.s_panel) cudaStreamDestroy(r.s_panel);
  if (r.s_comm) cudaStreamDestroy(r.s_comm);
  if (r.s_chain) cudaStreamDestroy(r.s_chain);
  smpart_destroy(r.part);
  r.part = nullptr;
}

// Cholesky of the diagonal block (k, k) in place + its inverse into Wb: the single-GPU recursion on one nb x nb block
static int diag_factor(DistRank& r, int k, cudaStream_t stream = nullptr, double* Wout = nullptr) {
  const int64_t gb = (int64_t)k * r.nb;
  if (!Wout) Wout = r.Wb;
  Dense d;
  d.s = stream ? stream : r.s_panel;
  d.launches = &r.launches;
  d.Dinv = nullptr;
  d.info = r.info;
  d.logdet = r.scal + DS_LOGDET;
  d.W = nullptr;
  d.nvalid = r.N;
  d.ldw = r.nb;
  d.Winv = Wout - (gb + gb * d.ldw);  // potrf_inv_rec addresses W by the block's global position
  d.tmpL = r.tmpL;
  d.Tpool = r.Tpool;
  d.TLpool = nullptr;
  double* Tkk = r.T + (int64_t)(k / r.P) * r.nb + (int64_t)(k / r.Q) * r.nb * r.ML;
  return potrf_inv_rec(d, Tkk, r.ML, r.nb, gb, d.Tpool, false, 0);
}

// K -> K^-1 in place (see the file header).  Everything is queued; the caller synchronises.
struct SweepProf {  // GPC_DIST_PROFILE=1: per-stage durations of the sweep (timing events on the stage's own stream)
  std::vector<cudaEvent_t> ev;
  int nbt = 0;
  cudaEvent_t at(int k, int i) { return ev[(size_t)k * 8 + i]; }
};
static SweepProf* sweep_prof(int nbt) {
  static thread_local SweepProf* p = nullptr;
  static int on = -1;
  if (on < 0) on = getenv("GPC_DIST_PROFILE") ? 1 : 0;
  if (!on) return nullptr;
  if (!p) p = new SweepProf();
  if (p->nbt < nbt) {
    size_t old = p->ev.size();
    p->ev.resize((size_t)nbt * 8);
    for (size_t i = old; i < p->ev.size(); i++) cudaEventCreate(&p->ev[i]);
    p->nbt = nbt;
  }
  return p;
}
static void sweep_prof_report(SweepProf* pf, int nbt) {
  if (!pf) return;
  const char* names[] = {"lookahead strips", "diag factor", "W bcast wait", "panel products + slicing", "slot bcast", "bulk update"};
  const int a[] = {0, 1, 2, 3, 4, 6}, b[] = {1, 2, 3, 4, 5, 7};
  double chain = 0.0;
  for (int st = 0; st < 6; st++) {
    double tot = 0.0;
    for (int k = 0; k < nbt; k++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, pf->at(k, a[st]), pf->at(k, b[st])) == cudaSuccess) tot += ms;
    }
    fprintf(stderr, "[gpc dist profile] %-26s total %8.3f ms  (%.3f ms / step)\n", names[st], tot, tot / nbt);
    if (st < 5) chain += tot;
  }
  fprintf(stderr, "[gpc dist profile] chain (panel + comm streams) %.3f ms, %d steps\n", chain, nbt);
}

static int sweep(DistRank& r) {
  const int nb = r.nb, NBt = r.NBt, P = r.P, Q = r.Q, p = r.p, q = r.q;
  SweepProf* pf = (r.rank == 0) ? sweep_prof(NBt) : nullptr;
#define PROF(k, i, stream) \
  if (pf) GPC_CUDA_CHECK(cudaEventRecord(pf->at(k, i), stream))
  const OzCycGrid gr{P, Q, p, q};
  const size_t sb = oz_slot_bytes(nb, r.S);
  std::vector<BcastItem> items((size_t)NBt);
  auto EV = [&](int kind, int k) { return r.ev[(size_t)kind * NBt + k]; };  // 0 slots arrived, 1 bulk done, 2 W ready, 3 W arrived, 4 panel produced
  for (int k = 0; k < NBt; k++) {
    const int b = k & 1;
    const bool col_owner = (q == k % Q), row_owner = (p == k % P), diag_owner = col_owner && row_owner;
    // ---- look-ahead: block column k and block row k receive the update of panel k-1 ahead of the bulk
    if (k >= 1) GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_panel, EV(0, k - 1), 0));
    if (k >= 2) GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_panel, EV(1, k - 2), 0));
    PROF(k, 0, r.s_panel);
    if (k >= 1) {
      if (col_owner) {
        const int64_t r0 = (int64_t)r.lrow_lb(k) * nb;
        GPC_CHECK(launch_oz_cyc_update(r.maps[(k - 1) & 1], gr, k - 1, -1, -1, r.T, r.ML, r0, r.ML - r0,
                                       (int64_t)(k / Q) * nb, nb, r.errflag, r.s_panel, &r.launches));
      }
      if (row_owner) {
        const int64_t n = (int64_t)r.lcol_lb(k) * nb;  // block columns with global index < k
        GPC_CHECK(launch_oz_cyc_update(r.maps[(k - 1) & 1], gr, k - 1, -1, -1, r.T, r.ML, (int64_t)(k / P) * nb, nb, 0, n,
                                       r.errflag, r.s_panel, &r.launches));
      }
    }
    // ---- diagonal block: L_kk (in place) and W_kk = L_kk^-1, then W_kk to everybody
    PROF(k, 1, r.s_panel);
    if (diag_owner) GPC_CHECK(diag_factor(r, k));
    PROF(k, 2, r.s_panel);
    GPC_CUDA_CHECK(cudaEventRecord(EV(2, k), r.s_panel));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_comm, EV(2, k), 0));
    {
      BcastItem w{r.Wb, (size_t)nb * nb * sizeof(double), r.owner(k, k)};
      GPC_CHECK(r.comm->bcast_group(&w, 1, r.s_comm));
    }
    GPC_CUDA_CHECK(cudaEventRecord(EV(3, k), r.s_comm));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_panel, EV(3, k), 0));
    PROF(k, 3, r.s_panel);
    // ---- this rank's blocks of panel k, sliced straight into their slots
    if (col_owner) {  // L_ik = A_ik W_kk'  for the local block rows with global index > k
      const int il0 = r.lrow_lb(k + 1);
      const int64_t m = r.ML - (int64_t)il0 * nb;
      if (m > 0) {
        GemmCall g{r.T + (int64_t)il0 * nb + (int64_t)(k / Q) * nb * r.ML, r.Wb, r.Lp, r.ML, nb, m, m, nb, nb, 1.0, 0.0,
                   false, false, false};
        g.b_tri = -1;
        GPC_CHECK(launch_gemm(g, r.s_panel, &r.launches));
        GPC_CHECK(oz_slice_to_slots(r.Lp, m, false, m, nb, r.S, r.emax, r.slots[b], il0 * P + p, P, r.s_panel, &r.launches));
      }
    }
    if (row_owner) {  // R_kj' = B_kj' W_kk'  for the local block columns with global index < k
      const int64_t m = (int64_t)r.lcol_lb(k) * nb;
      if (m > 0) {
        GemmCall g{r.T + (int64_t)(k / P) * nb, r.Wb, r.Lp, r.ML, nb, m, m, nb, nb, 1.0, 0.0, true, false, false};
        g.b_tri = -1;
        GPC_CHECK(launch_gemm(g, r.s_panel, &r.launches));
        GPC_CHECK(oz_slice_to_slots(r.Lp, m, false, m, nb, r.S, r.emax, r.slots[b], q, Q, r.s_panel, &r.launches));
      }
    }
    if (diag_owner)  // S_k = W_kk'
      GPC_CHECK(oz_slice_to_slots(r.Wb, nb, true, nb, nb, r.S, r.emax, r.slots[b], k, 0, r.s_panel, &r.launches));
    PROF(k, 4, r.s_panel);
    GPC_CUDA_CHECK(cudaEventRecord(EV(4, k), r.s_panel));
    // ---- the panel to everybody: one broadcast per slot, from the rank that produced it, all in ONE group.  (Sending it
    //      in chunks as they are sliced, so that transfer overlaps production, was measured on 8 GPUs and is slower: every
    //      additional NCCL group costs more launch latency than the overlap returns -- C3 0.092 -> 0.102 s.)
    GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_comm, EV(4, k), 0));
    for (int g = 0; g < NBt; g++) items[(size_t)g] = BcastItem{r.slots[b] + (size_t)g * sb, sb, r.producer(g, k)};
    GPC_CHECK(r.comm->bcast_group(items.data(), NBt, r.s_comm));
    PROF(k, 5, r.s_comm);
    GPC_CUDA_CHECK(cudaEventRecord(EV(0, k), r.s_comm));
    // ---- bulk update of step k: every local block except block row / column k+1 (done by the look-ahead of step k+1)
    GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_main, EV(0, k), 0));
    const int skip = (k + 1 < NBt) ? k + 1 : -1;
    PROF(k, 6, r.s_main);
    GPC_CHECK(launch_oz_cyc_update(r.maps[b], gr, k, skip, skip, r.T, r.ML, 0, r.ML, 0, r.NL, r.errflag, r.s_main,
                                   &r.launches));
    PROF(k, 7, r.s_main);
    GPC_CUDA_CHECK(cudaEventRecord(EV(1, k), r.s_main));
  }
  // the other two streams have nothing queued beyond what the main stream already waited for, except the last panel's
  // producers: join them
  cudaEvent_t e = r.ev[(size_t)8 * NBt];
  GPC_CUDA_CHECK(cudaEventRecord(e, r.s_panel));
  GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_main, e, 0));
  e = r.ev[(size_t)8 * NBt + 1];
  GPC_CUDA_CHECK(cudaEventRecord(e, r.s_comm));
  GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_main, e, 0));
#undef PROF
  if (pf) {
    GPC_CUDA_CHECK(cudaStreamSynchronize(r.s_main));
    sweep_prof_report(pf, NBt);
  }
  return GPC_OK;
}

// Single rank: the same sweep with the critical path cut down to what it really is.  The next diagonal block needs only
// ONE block of the current panel -- T_{k+1,k+1} -= L_{k+1,k} L_{k+1,k}' -- so a dedicated chain stream does
//     T_kk -= L_{k,k-1} L_{k,k-1}'  ->  potrf + inverse of T_kk  ->  L_{k+1,k} = T_{k+1,k} W_kk'        (fp64, one block each)
// while the rest of panel k (all block rows, slicing), the look-ahead strips and the bulk update run behind it on the
// panel / main streams.  (With several ranks block (k+1, k) and block (k+1, k+1) live on different GPUs in general.)
static int sweep_solo(DistRank& r) {
  const int nb = r.nb, NBt = r.NBt;
  const OzCycGrid gr{1, 1, 0, 0};
  // events: 0 slots ready, 1 step k applied to block row / column k+2 (second look-ahead, first thing on the main stream),
  //         2 W_kk ready, 3 look-ahead strips done, 4 L_{k+1,k} ready, 5 bulk update of step k done
  auto EV = [&](int kind, int k) { return r.ev[(size_t)kind * NBt + k]; };
  SweepProf* pf = sweep_prof(NBt);
#define PROF(k, i, stream) \
  if (pf) GPC_CUDA_CHECK(cudaEventRecord(pf->at(k, i), stream))
  {  // the chain stream starts after whatever the panel stream was told to wait for (the K build)
    cudaEvent_t e = r.ev[(size_t)8 * NBt + 3];
    GPC_CUDA_CHECK(cudaEventRecord(e, r.s_panel));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_chain, e, 0));
  }
  for (int k = 0; k < NBt; k++) {
    const int b = k & 1;
    double* Wk = b ? r.Wb2 : r.Wb;
    double* Tkk = r.T + (int64_t)k * nb + (int64_t)k * nb * r.ML;
    // ---- panel stream: panel k-1 applied to block column k (rows > k) and block row k (columns < k)
    PROF(k, 0, r.s_panel);
    if (k >= 1) {
      GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_panel, EV(0, k - 1), 0));
      if (k >= 2) GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_panel, EV(1, k - 2), 0));
      GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_panel, EV(4, k - 1), 0));  // the chain has read T_{k,k-1}
      const int64_t r0 = (int64_t)(k + 1) * nb;
      GPC_CHECK(launch_oz_cyc_update(r.maps[(k - 1) & 1], gr, k - 1, -1, -1, r.T, r.ML, r0, r.ML - r0, (int64_t)k * nb, nb,
                                     r.errflag, r.s_panel, &r.launches));
      GPC_CHECK(launch_oz_cyc_update(r.maps[(k - 1) & 1], gr, k - 1, -1, -1, r.T, r.ML, (int64_t)k * nb, nb, 0,
                                     (int64_t)k * nb, r.errflag, r.s_panel, &r.launches));
    }
    GPC_CUDA_CHECK(cudaEventRecord(EV(3, k), r.s_panel));
    PROF(k, 1, r.s_panel);
    // ---- chain stream: the diagonal block
    if (k >= 2) GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_chain, EV(1, k - 2), 0));
    PROF(k, 2, r.s_chain);
    if (k >= 1) {
      double* Lf = r.LF[(k - 1) & 1];
      GemmCall g{Lf, Lf, Tkk, nb, nb, r.ML, nb, nb, nb, -1.0, 1.0, false, false, true};
      GPC_CHECK(launch_gemm(g, r.s_chain, &r.launches));
    }
    GPC_CHECK(diag_factor(r, k, r.s_chain, Wk));
    GPC_CUDA_CHECK(cudaEventRecord(EV(2, k), r.s_chain));
    PROF(k, 3, r.s_chain);
    if (k + 1 < NBt) {  // L_{k+1,k} for the next diagonal block
      GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_chain, EV(3, k), 0));
      GemmCall g{r.T + (int64_t)(k + 1) * nb + (int64_t)k * nb * r.ML, Wk, r.LF[b], r.ML, nb, nb, nb, nb, nb, 1.0, 0.0, false,
                 false, false};
      g.b_tri = -1;
      GPC_CHECK(launch_gemm(g, r.s_chain, &r.launches));
    }
    GPC_CUDA_CHECK(cudaEventRecord(EV(4, k), r.s_chain));
    PROF(k, 4, r.s_chain);
    // ---- panel stream: the whole panel k, sliced into its slots (free once the bulk update of step k-2 has read them)
    GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_panel, EV(2, k), 0));
    if (k >= 2) GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_panel, EV(5, k - 2), 0));
    PROF(k, 5, r.s_panel);
    {
      const int64_t m = r.ML - (int64_t)(k + 1) * nb;
      if (m > 0) {
        GemmCall g{r.T + (int64_t)(k + 1) * nb + (int64_t)k * nb * r.ML, Wk, r.Lp, r.ML, nb, m, m, nb, nb, 1.0, 0.0, false, false,
                   false};
        g.b_tri = -1;
        GPC_CHECK(launch_gemm(g, r.s_panel, &r.launches));
        GPC_CHECK(oz_slice_to_slots(r.Lp, m, false, m, nb, r.S, r.emax, r.slots[b], k + 1, 1, r.s_panel, &r.launches));
      }
      const int64_t mr = (int64_t)k * nb;
      if (mr > 0) {
        GemmCall g{r.T + (int64_t)k * nb, Wk, r.Lp, r.ML, nb, mr, mr, nb, nb, 1.0, 0.0, true, false, false};
        g.b_tri = -1;
        GPC_CHECK(launch_gemm(g, r.s_panel, &r.launches));
        GPC_CHECK(oz_slice_to_slots(r.Lp, mr, false, mr, nb, r.S, r.emax, r.slots[b], 0, 1, r.s_panel, &r.launches));
      }
      GPC_CHECK(oz_slice_to_slots(Wk, nb, true, nb, nb, r.S, r.emax, r.slots[b], k, 0, r.s_panel, &r.launches));
    }
    GPC_CUDA_CHECK(cudaEventRecord(EV(0, k), r.s_panel));
    PROF(k, 6, r.s_panel);
    // ---- main stream: bulk update of step k (everything but block row / column k+1), after the chain has read T_{k+1,k}
    GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_main, EV(0, k), 0));
    GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_main, EV(4, k), 0));
    if (k + 2 < NBt) {  // second look-ahead: block column k+2 (rows >= k+2) and block row k+2 (columns <= k) first ...
      const int64_t c2 = (int64_t)(k + 2) * nb;
      GPC_CHECK(launch_oz_cyc_update(r.maps[b], gr, k, -1, -1, r.T, r.ML, c2, r.ML - c2, c2, nb, r.errflag, r.s_main,
                                     &r.launches));
      GPC_CHECK(launch_oz_cyc_update(r.maps[b], gr, k, -1, -1, r.T, r.ML, c2, nb, 0, (int64_t)(k + 1) * nb, r.errflag, r.s_main,
                                     &r.launches));
    }
    GPC_CUDA_CHECK(cudaEventRecord(EV(1, k), r.s_main));
    {  // ... then everything else (block rows / columns k+1 and k+2 left out)
      const int lo = (k + 1 < NBt) ? k + 1 : -1, hi = (k + 2 < NBt) ? k + 2 : k + 1;
      GPC_CHECK(launch_oz_cyc_update(r.maps[b], gr, k, lo, hi, r.T, r.ML, 0, r.ML, 0, r.NL, r.errflag, r.s_main, &r.launches));
    }
    GPC_CUDA_CHECK(cudaEventRecord(EV(5, k), r.s_main));
  }
  cudaEvent_t e = r.ev[(size_t)8 * NBt];
  GPC_CUDA_CHECK(cudaEventRecord(e, r.s_panel));
  GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_main, e, 0));
  e = r.ev[(size_t)8 * NBt + 1];
  GPC_CUDA_CHECK(cudaEventRecord(e, r.s_chain));
  GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_main, e, 0));
#undef PROF
  if (pf) {
    GPC_CUDA_CHECK(cudaStreamSynchronize(r.s_main));
    const char* names[] = {"look-ahead strips", "chain: SYRK + diag factor", "chain: L(k+1,k)", "panel products + slicing"};
    const int a[] = {0, 2, 3, 5}, bb[] = {1, 3, 4, 6};
    for (int st = 0; st < 4; st++) {
      double tot = 0.0;
      for (int k = 0; k < NBt; k++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pf->at(k, a[st]), pf->at(k, bb[st])) == cudaSuccess) tot += ms;
      }
      fprintf(stderr, "[gpc sweep profile] %-28s total %8.3f ms  (%.3f ms / step)\n", names[st], tot, tot / NBt);
    }
  }
  return GPC_OK;
}

struct EvalOut {
  double logdet = 0, quad = 0, jitter = 0;
  int info = 0;
  std::vector<double> g;
};

// one full evaluation on this rank (collective: every rank runs it)
static int rank_eval(DistRank& r, const KSpec& ks, EvalOut* out) {
  GPC_CUDA_CHECK(cudaSetDevice(r.device));
  if (!r.haveData) {
    set_error("gpc_dist_eval: X and m have not been set");
    return GPC_ERR_STATE;
  }
  cudaStream_t s = r.s_main;
  const CycMap cm = r.cyc();
  r.haveInv = false;
  double jitter = 0.0, jitter_used = 0.0;
  const int nsc = DS_G + ks.nparams;
  for (int tries = 0;; tries++) {
    GPC_CUDA_CHECK(cudaMemsetAsync(r.info, 0, sizeof(int), s));
    GPC_CUDA_CHECK(cudaMemsetAsync(r.scal, 0, (size_t)nsc * sizeof(double), s));
    GPC_CUDA_CHECK(cudaEventRecord(r.tev[0], s));
    // K: the local blocks on / below the diagonal (jitter of the retry schedule already on the diagonal)
    GPC_CHECK(launch_kbuild_cyc(ks, r.X, r.Np, r.N, r.T, r.ML, cm, jitter_used, s, &r.launches));
    GPC_CUDA_CHECK(cudaEventRecord(r.tev[1], s));
    {  // the panel stream starts after the K build
      cudaEvent_t e = r.ev[(size_t)8 * r.NBt + 2];
      GPC_CUDA_CHECK(cudaEventRecord(e, s));
      GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_panel, e, 0));
      GPC_CUDA_CHECK(cudaStreamWaitEvent(r.s_comm, e, 0));
    }
    if (r.world == 1 && !getenv("GPC_SWEEP_GENERIC")) GPC_CHECK(sweep_solo(r));
    else GPC_CHECK(sweep(r));
    GPC_CUDA_CHECK(cudaEventRecord(r.tev[2], s));
    // alpha = K^-1 m: local symmetric block product, all-reduce
    GPC_CUDA_CHECK(cudaMemsetAsync(r.y, 0, (size_t)r.Np * r.d * sizeof(double), s));
    GPC_CHECK(launch_symv_cyc(r.T, r.ML, cm, r.M, r.Np, r.d, r.y, r.Np, s, &r.launches));
    GPC_CHECK(r.comm->allreduce(r.y, (int)(r.Np * r.d), false, s));
    GPC_CUDA_CHECK(cudaMemcpyAsync(r.alpha, r.y, (size_t)r.Np * r.d * sizeof(double), cudaMemcpyDeviceToDevice, s));
    GPC_CUDA_CHECK(cudaEventRecord(r.tev[3], s));
    // gradient partial sums over the local blocks -> scal[DS_G..]
    GPC_CHECK(launch_grad(ks, r.X, r.Np, r.N, r.Np, r.T, r.ML, r.alpha, r.Np, r.d, 0, r.partial, r.max_ctas, r.scal + DS_G,
                          nullptr, 0, s, &r.launches, -1, 0, &cm));
    GPC_CUDA_CHECK(cudaEventRecord(r.tev[4], s));
    // log det (held by the diagonal owners) and the gradient: one all-reduce; first bad pivot: max
    GPC_CUDA_CHECK(cudaMemsetAsync(r.scal + DS_QUAD, 0, sizeof(double), s));
    GPC_CHECK(r.comm->allreduce(r.scal, nsc, false, s));
    GPC_CUDA_CHECK(cudaMemsetAsync(r.scal + DS_QUAD, 0, sizeof(double), s));
    GPC_CHECK(launch_dot(r.M, r.alpha, r.Np * r.d, r.scal + DS_QUAD, s, &r.launches));
    {  // info as a double for the max-reduction (scal[DS_INFO] is not part of the sum's payload of interest)
      GPC_CUDA_CHECK(cudaMemcpyAsync(r.hres + DS_INFO, r.info, sizeof(int), cudaMemcpyDeviceToHost, s));
    }
    GPC_CUDA_CHECK(cudaMemcpyAsync(r.hres, r.scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    GPC_CUDA_CHECK(cudaMemcpyAsync(r.hres + DS_G, r.scal + DS_G, (size_t)ks.nparams * sizeof(double), cudaMemcpyDeviceToHost, s));
    GPC_CUDA_CHECK(cudaEventRecord(r.tev[5], s));
    GPC_CUDA_CHECK(cudaStreamSynchronize(s));
    int myinfo = 0;
    memcpy(&myinfo, r.hres + DS_INFO, sizeof(int));
    // agree on the status: the largest "first non-positive pivot" over the ranks (0 everywhere = success)
    double* dinfo = r.scal + DS_INFO;
    double hinfo = (double)myinfo;
    GPC_CUDA_CHECK(cudaMemcpyAsync(dinfo, &hinfo, sizeof(double), cudaMemcpyHostToDevice, s));
    GPC_CHECK(r.comm->allreduce(dinfo, 1, true, s));
    GPC_CUDA_CHECK(cudaMemcpyAsync(&hinfo, dinfo, sizeof(double), cudaMemcpyDeviceToHost, s));
    GPC_CUDA_CHECK(cudaStreamSynchronize(s));
    const int info = (int)hinfo;
    if (info == 0) break;
    // jitChol schedule (CMatrix.cpp:767-804): 1e-6 * mean(diag K), x10 per retry; K is rebuilt with the accumulated jitter
    if (tries == 0) {
      // mean of the diagonal = sum of the components' diagComputeElement over the data: all ranks hold X, evaluate locally
      double* kd = r.y;  // scratch
      GPC_CHECK(launch_kdiag(ks, r.X, r.Np, r.N, kd, s, &r.launches));
      std::vector<double> h((size_t)r.N);
      GPC_CUDA_CHECK(cudaMemcpyAsync(h.data(), kd, (size_t)r.N * sizeof(double), cudaMemcpyDeviceToHost, s));
      GPC_CUDA_CHECK(cudaStreamSynchronize(s));
      double tr = 0.0;
      for (double v : h) tr += v;
      jitter = 1e-6 * tr / (double)r.N;
    }
    jitter_used += jitter;
    jitter *= 10.0;
    if (jitter > 10.0 || tries + 1 >= 20) {
      set_error("gpc_dist_eval: kernel matrix is non positive definite after jitter retries");
      out->info = info;
      return info;
    }
  }
  r.haveInv = true;
  out->logdet = r.hres[DS_LOGDET];
  out->quad = r.hres[DS_QUAD];
  out->jitter = jitter_used;
  out->info = 0;
  out->g.assign(r.hres + DS_G, r.hres + DS_G + ks.nparams);
  for (int i = 0; i < 5; i++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.tev[i], r.tev[i + 1]);
    r.last_ms[i] = ms;
  }
  int herr = 0;
  GPC_CUDA_CHECK(cudaMemcpy(&herr, r.errflag, sizeof(int), cudaMemcpyDeviceToHost));
  if (herr) {
    set_error("tensor-core GEMM pipeline protocol error " + std::to_string(herr));
    return GPC_ERR_CUDA;
  }
  return GPC_OK;
}

}  // namespace gpc

using namespace gpc;

// ------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------
struct gpc_dist {
  int mode = 0;  // 0 local (threads), 1 nccl (this process = one rank)
  int P = 1, Q = 1, world = 1;
  std::vector<DistRank> ranks;  // local: world entries; nccl: one
  std::vector<Comm*> comms;
  LocalHub hub;
};

static void grid_setup(DistRank& r, int rank, int world, int P, int Q, int64_t N, int D, int dout, int nb) {
  r.rank = rank;
  r.world = world;
  r.P = P;
  r.Q = Q;
  r.p = rank / Q;
  r.q = rank % Q;
  r.N = N;
  r.nb = nb;
  r.Np = round_up(N, nb);
  r.NBt = (int)(r.Np / nb);
  r.D = D;
  r.d = dout;
  r.S = oz_slices();
  r.nlr = r.NBt > r.p ? (r.NBt - r.p + P - 1) / P : 0;
  r.nlc = r.NBt > r.q ? (r.NBt - r.q + Q - 1) / Q : 0;
  r.ML = (int64_t)r.nlr * nb;
  r.NL = (int64_t)r.nlc * nb;
}

static int check_shape(int world, int P, int Q, int64_t N, int D, int dout, int nb) {
  if (world < 1 || P < 1 || Q < 1 || P * Q != world || N < 1 || D < 1 || dout < 1 || nb < 128 || nb % 128 || nb > 16384) {
    set_error("gpc_dist_create: need P * Q == number of ranks, nb a multiple of 128 (<= 16384)");
    return GPC_ERR_ARG;
  }
  if (round_up(N, nb) / nb > 4096) {
    set_error("gpc_dist_create: more than 4096 block rows; use a larger nb");
    return GPC_ERR_ARG;
  }
  return GPC_OK;
}

template <class F>
static int for_each_rank(gpc_dist* h, F fn) {
  if (h->ranks.size() == 1) return fn(h->ranks[0]);
  std::vector<std::thread> th;
  for (auto& r : h->ranks)
    th.emplace_back([&h, &r, &fn]() {
      cudaSetDevice(r.device);
      r.result = fn(r);
      if (r.result != GPC_OK) {
        r.err = gpc_last_error();
        if (r.result < 0) h->hub.fail();  // nobody may wait for this rank any more
      }
    });
  for (auto& t : th) t.join();
  int rc = GPC_OK;
  for (auto& r : h->ranks)
    if (r.result != GPC_OK && rc == GPC_OK) {
      rc = r.result;
      set_error("rank " + std::to_string(r.rank) + ": " + r.err);
    }
  return rc;
}

extern "C" {

int gpc_dist_unique_id(void* id128) {
  if (!id128) return GPC_ERR_ARG;
  NcclApi* api = nccl_api();
  if (!api) return GPC_ERR_CUDA;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  GPC_NCCL_CHECK(api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return GPC_OK;
}

int gpc_dist_destroy(gpc_dist* h) {
  if (!h) return GPC_OK;
  for (auto& r : h->ranks) rank_free(r);
  for (Comm* c : h->comms) delete c;
  for (auto e : h->hub.e_ready)
    if (e) cudaEventDestroy(e);
  for (auto e : h->hub.e_done)
    if (e) cudaEventDestroy(e);
  for (auto p : h->hub.host)
    if (p) cudaFreeHost(p);
  delete h;
  return GPC_OK;
}

int gpc_dist_create_nccl(gpc_dist** out, int device, int rank, int world, const void* id128, int P, int Q, int64_t N, int D,
                         int dout, int nb) {
  if (!out || rank < 0 || rank >= world || (world > 1 && !id128)) {
    set_error("gpc_dist_create_nccl: bad arguments");
    return GPC_ERR_ARG;
  }
  *out = nullptr;
  GPC_CHECK(check_shape(world, P, Q, N, D, dout, nb));
  int ndev = 0;
  GPC_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) {
    set_error("gpc_dist_create_nccl: no such CUDA device");
    return GPC_ERR_CUDA;
  }
  GPC_CUDA_CHECK(cudaSetDevice(device));
  gpc_dist* h = new gpc_dist();
  h->mode = 1;
  h->P = P;
  h->Q = Q;
  h->world = world;
  h->ranks.resize(1);
  DistRank& r = h->ranks[0];
  r.device = device;
  grid_setup(r, rank, world, P, Q, N, D, dout, nb);
  int rc = GPC_OK;
  if (world == 1) {
    h->comms.push_back(new SoloComm());
  } else {
    NcclApi* api = nccl_api();
    if (!api) {
      gpc_dist_destroy(h);
      return GPC_ERR_CUDA;
    }
    NcclComm* c = new NcclComm();
    c->api = api;
    h->comms.push_back(c);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t nr = api->CommInitRank(&c->comm, world, id, rank);
    if (nr != ncclSuccess) {
      set_error(std::string("ncclCommInitRank: ") + api->GetErrorString(nr));
      rc = GPC_ERR_CUDA;
    }
  }
  h->comms[0]->rank = rank;
  h->comms[0]->world = world;
  r.comm = h->comms[0];
  if (rc == GPC_OK) rc = rank_alloc(r);
  if (rc != GPC_OK) {
    std::string keep = gpc_last_error();
    gpc_dist_destroy(h);
    cudaGetLastError();
    set_error(keep);
    return rc;
  }
  *out = h;
  return GPC_OK;
}

int gpc_dist_create_local(gpc_dist** out, const int* devices, int ndev, int P, int Q, int64_t N, int D, int dout, int nb) {
  if (!out || !devices || ndev < 1) {
    set_error("gpc_dist_create_local: bad arguments");
    return GPC_ERR_ARG;
  }
  *out = nullptr;
  GPC_CHECK(check_shape(ndev, P, Q, N, D, dout, nb));
  int have = 0;
  GPC_CUDA_CHECK(cudaGetDeviceCount(&have));
  for (int i = 0; i < ndev; i++)
    if (devices[i] < 0 || devices[i] >= have) {
      set_error("gpc_dist_create_local: no such CUDA device");
      return GPC_ERR_CUDA;
    }
  gpc_dist* h = new gpc_dist();
  h->mode = 0;
  h->P = P;
  h->Q = Q;
  h->world = ndev;
  h->ranks.resize((size_t)ndev);
  LocalHub& hub = h->hub;
  hub.world = ndev;
  hub.dev.assign(devices, devices + ndev);
  hub.e_ready.assign((size_t)ndev, nullptr);
  hub.e_done.assign((size_t)ndev, nullptr);
  hub.ptrs.resize((size_t)ndev);
  hub.host.assign((size_t)ndev, nullptr);
  int rc = GPC_OK;
  for (int i = 0; i < ndev && rc == GPC_OK; i++) {
    DistRank& r = h->ranks[(size_t)i];
    r.device = devices[i];
    grid_setup(r, i, ndev, P, Q, N, D, dout, nb);
    Comm* c;
    if (ndev == 1) {
      c = new SoloComm();
    } else {
      LocalComm* lc = new LocalComm();
      lc->hub = &hub;
      c = lc;
    }
    c->rank = i;
    c->world = ndev;
    h->comms.push_back(c);
    r.comm = c;
    if (cudaSetDevice(devices[i]) != cudaSuccess) rc = GPC_ERR_CUDA;
    if (rc == GPC_OK && cudaEventCreateWithFlags(&hub.e_ready[(size_t)i], cudaEventDisableTiming) != cudaSuccess) rc = GPC_ERR_CUDA;
    if (rc == GPC_OK && cudaEventCreateWithFlags(&hub.e_done[(size_t)i], cudaEventDisableTiming) != cudaSuccess) rc = GPC_ERR_CUDA;
    hub.host_cap = (size_t)(r.Np * dout > DS_G + GPC_MAX_PARAMS ? r.Np * dout : DS_G + GPC_MAX_PARAMS);
    if (rc == GPC_OK && cudaMallocHost(&hub.host[(size_t)i], 2 * hub.host_cap * sizeof(double)) != cudaSuccess) rc = GPC_ERR_NOMEM;
    if (rc == GPC_OK) {
      for (int j = 0; j < ndev; j++)  // peer access for the direct copies (ignored when already enabled / same device)
        if (devices[j] != devices[i]) {
          int can = 0;
          cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
          if (can && cudaDeviceEnablePeerAccess(devices[j], 0) != cudaSuccess) cudaGetLastError();
        }
      rc = rank_alloc(r);
    } else {
      set_error("gpc_dist_create_local: CUDA set-up failed");
    }
  }
  if (rc != GPC_OK) {
    std::string keep = gpc_last_error();
    gpc_dist_destroy(h);
    cudaGetLastError();
    set_error(keep);
    return rc;
  }
  *out = h;
  return GPC_OK;
}

int gpc_dist_set_data(gpc_dist* h, const double* X, int64_t ldx, const double* M, int64_t ldm) {
  if (!h || !X || !M) return GPC_ERR_ARG;
  for (auto& r : h->ranks) {
    if (ldx < r.N || ldm < r.N) {
      set_error("gpc_dist_set_data: leading dimension smaller than N");
      return GPC_ERR_ARG;
    }
    GPC_CUDA_CHECK(cudaSetDevice(r.device));
    GPC_CUDA_CHECK(cudaMemsetAsync(r.X, 0, (size_t)r.Np * r.D * sizeof(double), r.s_main));
    GPC_CUDA_CHECK(cudaMemsetAsync(r.M, 0, (size_t)r.Np * r.d * sizeof(double), r.s_main));
    GPC_CUDA_CHECK(cudaMemcpy2DAsync(r.X, r.Np * sizeof(double), X, ldx * sizeof(double), r.N * sizeof(double), r.D,
                                     cudaMemcpyHostToDevice, r.s_main));
    GPC_CUDA_CHECK(cudaMemcpy2DAsync(r.M, r.Np * sizeof(double), M, ldm * sizeof(double), r.N * sizeof(double), r.d,
                                     cudaMemcpyHostToDevice, r.s_main));
    GPC_CUDA_CHECK(cudaStreamSynchronize(r.s_main));
    r.haveData = true;
    r.haveInv = false;
  }
  return GPC_OK;
}

int gpc_dist_eval(gpc_dist* h, const gpc_kcomp* comps, int ncomp, double* out, double* gparams) {
  if (!h) return GPC_ERR_ARG;
  KSpec ks;
  GPC_CHECK(make_kspec(comps, ncomp, h->ranks[0].D, &ks));
  std::vector<EvalOut> outs(h->ranks.size());
  int rc = for_each_rank(h, [&](DistRank& r) { return rank_eval(r, ks, &outs[(size_t)(&r - &h->ranks[0])]); });
  if (rc != GPC_OK) return rc;
  const EvalOut& o = outs[0];
  if (out) {
    out[0] = o.logdet;
    out[1] = o.quad;
    out[2] = o.jitter;
  }
  if (gparams)
    for (int i = 0; i < ks.nparams; i++) gparams[i] = o.g[(size_t)i];
  return GPC_OK;
}

int gpc_dist_download_kinv(gpc_dist* h, double* dst, int64_t ld) {
  if (!h || !dst) return GPC_ERR_ARG;
  for (auto& r : h->ranks) {
    if (!r.haveInv) {
      set_error("gpc_dist_download_kinv: no successful evaluation yet");
      return GPC_ERR_STATE;
    }
    if (ld < r.N) return GPC_ERR_ARG;
    if (r.ML == 0 || r.NL == 0) continue;
    GPC_CUDA_CHECK(cudaSetDevice(r.device));
    std::vector<double> loc((size_t)r.ML * (size_t)r.NL);
    GPC_CUDA_CHECK(cudaMemcpy(loc.data(), r.T, loc.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int jl = 0; jl < r.nlc; jl++)
      for (int il = 0; il < r.nlr; il++) {
        const int gi = il * r.P + r.p, gj = jl * r.Q + r.q;
        if (gi < gj) continue;
        for (int c = 0; c < r.nb; c++) {
          const int64_t j = (int64_t)gj * r.nb + c;
          if (j >= r.N) break;
          for (int rr = 0; rr < r.nb; rr++) {
            const int64_t i = (int64_t)gi * r.nb + rr;
            if (i >= r.N) break;
            const double v = loc[(size_t)(il * (int64_t)r.nb + rr) + (size_t)(jl * (int64_t)r.nb + c) * (size_t)r.ML];
            if (gi == gj && i < j) continue;  // diagonal blocks: take the lower half, mirror below
            dst[i + j * ld] = v;
            dst[j + i * ld] = v;
          }
        }
      }
  }
  return GPC_OK;
}

int gpc_dist_plan(int P, int Q, int rank, int64_t N, int nb, int k, int* out12, int* producers) {
  // host-only view of what `rank` does at step k (the helpers the sweep itself uses): CPU tests of the schedule
  if (P < 1 || Q < 1 || rank < 0 || rank >= P * Q || N < 1 || nb < 1 || !out12) return GPC_ERR_ARG;
  DistRank r;
  grid_setup(r, rank, P * Q, P, Q, N, 1, 1, nb);
  if (k < 0 || k >= r.NBt) return GPC_ERR_ARG;
  const bool col_owner = (r.q == k % Q), row_owner = (r.p == k % P);
  out12[0] = r.NBt;
  out12[1] = r.nlr;
  out12[2] = r.nlc;
  out12[3] = col_owner;
  out12[4] = row_owner;
  out12[5] = r.owner(k, k);
  // column part of the panel: local block rows [il0, nlr) -> slots il0 * P + p, stride P
  out12[6] = col_owner ? r.lrow_lb(k + 1) : -1;
  out12[7] = col_owner ? r.lrow_lb(k + 1) * P + r.p : -1;
  // row part: local block columns [0, cnt) -> slots q, q + Q, ...
  out12[8] = row_owner ? r.lcol_lb(k) : -1;
  // look-ahead strips of step k (panel k-1 applied to block column / row k): first local block row, number of block columns
  out12[9] = col_owner ? r.lrow_lb(k) : -1;
  out12[10] = row_owner ? r.lcol_lb(k) : -1;
  out12[11] = (k + 1 < r.NBt) ? k + 1 : -1;  // block row / column the bulk update of step k leaves out
  if (producers)
    for (int g = 0; g < r.NBt; g++) producers[g] = r.producer(g, k);
  return GPC_OK;
}

int gpc_dist_info(gpc_dist* h, int64_t* out8, double* ms5) {
  if (!h) return GPC_ERR_ARG;
  const DistRank& r = h->ranks[0];
  if (out8) {
    const size_t sb = oz_slot_bytes(r.nb, r.S);
    out8[0] = h->world;
    out8[1] = r.NBt;                                                    // steps
    out8[2] = (int64_t)r.ML * r.NL * 8;                                 // bytes of the local matrix (rank 0 / this rank)
    out8[3] = (int64_t)(2 * sb * (size_t)r.NBt);                        // bytes of the two panel buffers
    out8[4] = (int64_t)(sb * (size_t)r.NBt + (size_t)r.nb * r.nb * 8);  // bytes broadcast per step (panel + W_kk)
    int64_t l = 0;
    for (const auto& q : h->ranks) l += q.launches;
    out8[5] = l;
    out8[6] = r.nb;
    out8[7] = h->mode;
  }
  if (ms5)
    for (int i = 0; i < 5; i++) ms5[i] = r.last_ms[i];
  return GPC_OK;
}

}  // extern "C"
