// smpart.cu -- spatial partition of the SMs of one B200 into a small "chain" set and a large "bulk" set (CUDA green
// contexts, driver API resolved at run time through cudaGetDriverEntryPoint: the library does not link libcuda).
//
// Why: the Cholesky / inverse of one evaluation is a serial chain of small kernels (the 128 x 128 diagonal blocks and
// the 128..512-wide products between them, reference CMatrix.cpp:371-403 dpotrf_ inside jitChol) next to a few large
// tensor-core products.  The tensor-core kernel owns an SM completely (231 KB of shared memory, all 512 TMEM columns),
// so a chain kernel launched while a bulk product runs waits for a tile to drain on EVERY launch, whatever its stream
// priority.  With the SMs partitioned, the chain streams own their SMs and never queue behind a tile.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"

namespace gpc {

namespace {
struct DriverApi {
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                                        unsigned int) = nullptr;
  CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
  CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  bool ok = false;
};

template <class F>
bool resolve(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}

DriverApi* driver_api() {
  static DriverApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.ok = resolve("cuDeviceGet", &api.DeviceGet) && resolve("cuDeviceGetDevResource", &api.DeviceGetDevResource) &&
             resolve("cuDevSmResourceSplitByCount", &api.DevSmResourceSplitByCount) &&
             resolve("cuDevResourceGenerateDesc", &api.DevResourceGenerateDesc) &&
             resolve("cuGreenCtxCreate", &api.GreenCtxCreate) && resolve("cuGreenCtxDestroy", &api.GreenCtxDestroy) &&
             resolve("cuGreenCtxStreamCreate", &api.GreenCtxStreamCreate);
  }
  return api.ok ? &api : nullptr;
}
}  // namespace

struct SmPartition {
  CUgreenCtx chain = nullptr, bulk = nullptr;
  int sm_chain = 0, sm_bulk = 0;
};

// chain_sms <= 0: the environment decides (GPC_SM_PARTITION, default `dflt`; 0 = no partition).  Returns null when the
// partition is switched off or the driver cannot provide it; callers then fall back to ordinary priority streams.
SmPartition* smpart_create(int device, int chain_sms, int dflt) {
  if (chain_sms <= 0) {
    const char* e = getenv("GPC_SM_PARTITION");
    chain_sms = e ? atoi(e) : dflt;
  }
  if (chain_sms <= 0) return nullptr;
  DriverApi* api = driver_api();
  if (!api) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) {  // primary context up
    cudaGetLastError();
    return nullptr;
  }
  CUdevice dev;
  CUdevResource all, grp, rem;
  if (api->DeviceGet(&dev, device) != CUDA_SUCCESS) return nullptr;
  if (api->DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return nullptr;
  if ((unsigned)chain_sms + 8 > all.sm.smCount) return nullptr;
  unsigned int n = 1;
  if (api->DevSmResourceSplitByCount(&grp, &n, &all, &rem, 0, (unsigned)chain_sms) != CUDA_SUCCESS || n != 1) return nullptr;
  if (rem.sm.smCount == 0) return nullptr;
  SmPartition* p = new SmPartition();
  CUdevResourceDesc d1 = nullptr, d2 = nullptr;
  bool ok = api->DevResourceGenerateDesc(&d1, &grp, 1) == CUDA_SUCCESS &&
            api->GreenCtxCreate(&p->chain, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS &&
            api->DevResourceGenerateDesc(&d2, &rem, 1) == CUDA_SUCCESS &&
            api->GreenCtxCreate(&p->bulk, d2, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS;
  if (!ok) {
    if (p->chain) api->GreenCtxDestroy(p->chain);
    if (p->bulk) api->GreenCtxDestroy(p->bulk);
    delete p;
    return nullptr;
  }
  p->sm_chain = (int)grp.sm.smCount;
  p->sm_bulk = (int)rem.sm.smCount;
  if (getenv("GPC_TRACE"))
    fprintf(stderr, "[gpc] SM partition on device %d: %d chain + %d bulk SMs\n", device, p->sm_chain, p->sm_bulk);
  return p;
}

void smpart_destroy(SmPartition* p) {
  if (!p) return;
  DriverApi* api = driver_api();
  if (api) {
    if (p->chain) api->GreenCtxDestroy(p->chain);
    if (p->bulk) api->GreenCtxDestroy(p->bulk);
  }
  delete p;
}

int smpart_sms(const SmPartition* p, bool chain) { return p ? (chain ? p->sm_chain : p->sm_bulk) : 0; }

// a non-blocking stream confined to one side of the partition (destroy with cudaStreamDestroy)
int smpart_stream(SmPartition* p, bool chain, int priority, cudaStream_t* out) {
  DriverApi* api = driver_api();
  if (!p || !api) return GPC_ERR_STATE;
  CUstream s = nullptr;
  CUresult r = api->GreenCtxStreamCreate(&s, chain ? p->chain : p->bulk, CU_STREAM_NON_BLOCKING, priority);
  if (r != CUDA_SUCCESS) {
    set_error("cuGreenCtxStreamCreate failed: " + std::to_string((int)r));
    return GPC_ERR_CUDA;
  }
  *out = (cudaStream_t)s;
  return GPC_OK;
}

}  // namespace gpc
