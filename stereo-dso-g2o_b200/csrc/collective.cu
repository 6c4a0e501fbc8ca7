// The one exchange step of the path: point-sharded windowed BA sums the ranks' partial reduced camera
// systems [(4+8n)^2 + (4+8n) doubles, + the linearisation energy] with ONE NCCL allreduce over NVLink/NVSwitch, enqueued
// on the context's stream between ba_assemble_kernel and ba_solve_kernel. NCCL is resolved with dlopen at first use so the
// library carries no link-time dependency (single-GPU users never touch it); the communicator is created from a unique id
// that the host exchanges through whatever rendezvous it already has (bench.py: torch.distributed broadcast).
#include "ba_state.h"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>

namespace sdso {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api(std::string* err) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
    if (api.handle) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
      api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    }
  }
  if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce) {
    if (err) *err = "NCCL (libnccl.so.2) could not be loaded";
    return nullptr;
  }
  return &api;
}

void collective_destroy(sdso_ctx* ctx) {
  if (ctx->nccl_comm) {
    NcclApi* api = nccl_api(nullptr);
    if (api && api->CommDestroy) api->CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
}

}  // namespace sdso

using namespace sdso;

extern "C" {

// Contiguous block of the allPoints order owned by `rank` (partition by point so that each point's
// Hdd/bd/Hcd and its O(res^2) Schur terms stay rank-local). Blocks differ in size by at most one point.
int sdso_shard_range(int npoints, int rank, int nranks, int* begin, int* end) {
  if (npoints < 0 || nranks < 1 || rank < 0 || rank >= nranks || !begin || !end) return SDSO_E_INVALID;
  const int base = npoints / nranks, rem = npoints % nranks;
  *begin = rank * base + (rank < rem ? rank : rem);
  *end = *begin + base + (rank < rem ? 1 : 0);
  return SDSO_OK;
}

int sdso_nccl_unique_id(unsigned char id[128]) {
  if (!id) return SDSO_E_INVALID;
  NcclApi* api = nccl_api(nullptr);
  if (!api) return SDSO_E_STATE;
  ncclUniqueId u;
  if (api->GetUniqueId(&u) != ncclSuccess) return SDSO_E_CUDA;
  memcpy(id, u.internal, 128);
  return SDSO_OK;
}

int sdso_nccl_init(sdso_ctx* ctx, int rank, int nranks, const unsigned char id[128]) {
  sdso::enter(ctx);
  if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return SDSO_E_INVALID;
  std::string err;
  NcclApi* api = nccl_api(&err);
  if (!api) return fail(ctx, SDSO_E_STATE, err);
  collective_destroy(ctx);
  SDSO_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  ncclComm_t comm;
  ncclResult_t r = api->CommInitRank(&comm, nranks, u, rank);
  if (r != ncclSuccess) return fail(ctx, SDSO_E_CUDA, std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"));
  ctx->nccl_comm = comm; ctx->nccl_rank = rank; ctx->nccl_nranks = nranks;
  return SDSO_OK;
}

int sdso_nccl_destroy(sdso_ctx* ctx) {
  sdso::enter(ctx);
  if (!ctx) return SDSO_E_INVALID;
  collective_destroy(ctx);
  return SDSO_OK;
}

// In-place sum over ranks of a device buffer of doubles on the context's stream (asynchronous).
int sdso_allreduce_f64(sdso_ctx* ctx, void* device_buffer, int count) {
  sdso::enter(ctx);
  if (!ctx || !device_buffer || count < 0) return SDSO_E_INVALID;
  if (!ctx->nccl_comm) return fail(ctx, SDSO_E_STATE, "sdso_nccl_init has not been called");
  NcclApi* api = nccl_api(nullptr);
  ncclResult_t r = api->AllReduce(device_buffer, device_buffer, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream);
  if (r != ncclSuccess) return fail(ctx, SDSO_E_CUDA, std::string("ncclAllReduce: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"));
  ctx->launches++;
  return SDSO_OK;
}

}  // extern "C"
