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

// ---- one-shot allreduce over peer memory -----------------------------------------------------------------------------------------
// The exchange of the sharded BA is 7 141 doubles (57 KB) per LM iteration: far below the size where a ring or a tree pays, and
// NCCL's launch + protocol latency is the whole cost (13 us on 2 GPUs, 99 us on 8, measured). Here every rank owns one
// cudaMalloc'ed exchange block that all its peers map through CUDA IPC:
//     [ 256 B header | slots[parity 0..1][source rank 0..world-1] ],   slot = max_doubles x 16 bytes
// and ONE kernel per rank pushes and reduces, with no fence, no atomic and no barrier: every 8-byte word that crosses the link
// carries 4 bytes of payload and the 4-byte epoch of the exchange (the "low latency" wire format: an aligned 8-byte store
// arrives whole, so a word whose flag equals the epoch IS that epoch's payload):
//   1. thread i splits its double into two such words and stores them into slot[epoch & 1][rank] of every PEER's block
//      (posted stores over NVLink: no round trip);
//   2. the same thread polls the words of element i in the sources' slots of its OWN block (local memory) until both flags
//      equal the epoch, and sums the sources in rank order (its own contribution from the register) — the same order on every
//      rank, so all ranks end up with bit-identical sums and the redundant solves that follow stay in lock step.
// Two parities: a rank can only start epoch e+2 (which overwrites the parity of e) after its epoch-(e+1) kernel has finished,
// i.e. after it received every peer's epoch-(e+1) words, which a peer sends only from the kernel that runs after its epoch-e
// kernel has finished reading. A wait that exceeds ~2 s of SM clocks gives up and raises `err` (a lost peer must not hang the box).
// Earlier versions, measured on 2 GPUs (NCCL: 13.5 us per exchange): peers READ each other's slots behind a flag, 22.7 us; push +
// system fence + last-CTA flag, 16.0 us.
struct PeerExchange {
  int rank = 0, world = 1;
  size_t slot_doubles = 0;
  unsigned char* local = nullptr;          // this rank's block
  unsigned char* base[16] = {nullptr};     // every rank's block in this process' address space (base[rank] == local)
  bool opened[16] = {false};
  unsigned long long epoch = 0;
  unsigned* d_arrive = nullptr;            // [1] error flag
  bool connected = false, enabled = true;
};

struct PeerPtrs { unsigned char* base[16]; };
constexpr size_t kPeerHeader = 256;
constexpr size_t kPeerWordBytes = 16;   // per double: two (payload32, epoch32) words

__global__ void __launch_bounds__(256) peer_allreduce_kernel(double* buf, int count, PeerPtrs P, int rank, int world, unsigned epoch32,
                                                             size_t slot_doubles, unsigned* arrive) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  const size_t slot_bytes = slot_doubles * kPeerWordBytes;
  const size_t par_off = kPeerHeader + (size_t)(epoch32 & 1u) * world * slot_bytes;
  const unsigned long long tag = (unsigned long long)epoch32 << 32;
  for (int i0 = gtid; i0 < count; i0 += gthreads * 2) {
    // two elements per thread and round: their stores go out back to back, their polls overlap
    const int i1 = i0 + gthreads;
    const bool two = i1 < count;
    const double v0 = buf[i0], v1 = two ? buf[i1] : 0.0;
    const unsigned long long b0 = (unsigned long long)__double_as_longlong(v0), b1 = (unsigned long long)__double_as_longlong(v1);
    const ulonglong2 w0 = make_ulonglong2((b0 & 0xffffffffull) | tag, (b0 >> 32) | tag);
    const ulonglong2 w1 = make_ulonglong2((b1 & 0xffffffffull) | tag, (b1 >> 32) | tag);
    for (int r = 0; r < world; r++) {
      if (r == rank) continue;
      ulonglong2* dst = reinterpret_cast<ulonglong2*>(P.base[r] + par_off + (size_t)rank * slot_bytes);
      // volatile: the stores must be issued before the polling loop below, whatever the optimiser thinks of their addresses
      asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + i0), "l"(w0.x), "l"(w0.y) : "memory");
      if (two) asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + i1), "l"(w1.x), "l"(w1.y) : "memory");
    }
    double s0 = 0.0, s1 = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < world; r++) {
      if (r == rank) { s0 += v0; s1 += v1; continue; }
      const volatile ulonglong2* src = reinterpret_cast<const volatile ulonglong2*>(P.base[rank] + par_off + (size_t)r * slot_bytes);
      unsigned long long a, c;
      bool got0 = false, got1 = !two;
      unsigned long long lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
      while (!(got0 && got1)) {
        if (!got0) { a = src[i0].x; c = src[i0].y; if ((a >> 32) == epoch32 && (c >> 32) == epoch32) { lo0 = a; hi0 = c; got0 = true; } }
        if (!got1) { a = src[i1].x; c = src[i1].y; if ((a >> 32) == epoch32 && (c >> 32) == epoch32) { lo1 = a; hi1 = c; got1 = true; } }
        if (!(got0 && got1) && clock64() - t0 > 4000000000ll) { arrive[1] = 1; break; }
      }
      s0 += __longlong_as_double((long long)((lo0 & 0xffffffffull) | (hi0 << 32)));
      if (two) s1 += __longlong_as_double((long long)((lo1 & 0xffffffffull) | (hi1 << 32)));
    }
    buf[i0] = s0;
    if (two) buf[i1] = s1;
  }
}

static void peer_destroy(sdso_ctx* ctx) {
  PeerExchange* p = ctx->peer;
  if (!p) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < 16; r++) if (p->opened[r] && p->base[r]) cudaIpcCloseMemHandle(p->base[r]);
  if (p->local) cudaFree(p->local);
  if (p->d_arrive) cudaFree(p->d_arrive);
  delete p;
  ctx->peer = nullptr;
}

void collective_destroy(sdso_ctx* ctx) {
  peer_destroy(ctx);
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

// ---- peer-memory exchange: allocate + export, connect, select
int sdso_peer_alloc(sdso_ctx* ctx, int nranks, int max_doubles, unsigned char handle_out[64]) {
  sdso::enter(ctx);
  if (!ctx || nranks < 1 || nranks > 16 || max_doubles < 1 || !handle_out) return SDSO_E_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  if (ctx->peer) return fail(ctx, SDSO_E_STATE, "sdso_peer_alloc: the exchange block exists already");
  PeerExchange* p = new PeerExchange();
  p->slot_doubles = ((size_t)max_doubles + 31) & ~(size_t)31;
  p->world = nranks;
  const size_t bytes = kPeerHeader + 2 * (size_t)nranks * p->slot_doubles * kPeerWordBytes;
  if (cudaMalloc(&p->local, bytes) != cudaSuccess || cudaMalloc(&p->d_arrive, 2 * sizeof(unsigned)) != cudaSuccess) {
    if (p->local) cudaFree(p->local);
    delete p;
    cudaGetLastError();
    return fail(ctx, SDSO_E_NOMEM, "sdso_peer_alloc: cudaMalloc failed");
  }
  cudaMemset(p->local, 0, bytes);
  cudaMemset(p->d_arrive, 0, 2 * sizeof(unsigned));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p->local);
  if (e != cudaSuccess) { cudaFree(p->local); cudaFree(p->d_arrive); delete p; return fail(ctx, SDSO_E_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
  memcpy(handle_out, &h, 64);
  cudaDeviceSynchronize();
  ctx->peer = p;
  return SDSO_OK;
}

int sdso_peer_connect(sdso_ctx* ctx, int rank, int nranks, const unsigned char* handles /* nranks x 64 bytes, rank order */) {
  sdso::enter(ctx);
  if (!ctx || !handles || nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks) return SDSO_E_INVALID;
  PeerExchange* p = ctx->peer;
  if (!p) return fail(ctx, SDSO_E_STATE, "sdso_peer_connect before sdso_peer_alloc");
  if (nranks != p->world) return fail(ctx, SDSO_E_INVALID, "sdso_peer_connect: nranks differs from sdso_peer_alloc");
  p->rank = rank;
  for (int r = 0; r < nranks; r++) {
    if (r == rank) { p->base[r] = p->local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, 64);
    void* q = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(ctx, SDSO_E_CUDA, std::string("cudaIpcOpenMemHandle (peer access between the GPUs is required): ") + cudaGetErrorString(e));
    p->base[r] = static_cast<unsigned char*>(q); p->opened[r] = true;
  }
  p->connected = true;
  return SDSO_OK;
}

/* which = 0: NCCL, 1: the peer-memory kernel (default once connected) */
int sdso_peer_select(sdso_ctx* ctx, int which) {
  if (!ctx || !ctx->peer) return SDSO_E_INVALID;
  ctx->peer->enabled = which != 0;
  return SDSO_OK;
}

/* 0 = fine, 1 = a wait on a peer gave up (results of that exchange are not valid); synchronises the stream */
int sdso_peer_status(sdso_ctx* ctx, int* timed_out) {
  sdso::enter(ctx);
  if (!ctx || !ctx->peer || !timed_out) return SDSO_E_INVALID;
  unsigned v[2] = {0, 0};
  SDSO_CUDA(ctx, cudaMemcpyAsync(v, ctx->peer->d_arrive, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *timed_out = (int)v[1];
  return SDSO_OK;
}

// In-place sum over ranks of a device buffer of doubles on the context's stream (asynchronous).
int sdso_allreduce_f64(sdso_ctx* ctx, void* device_buffer, int count) {
  sdso::enter(ctx);
  if (!ctx || !device_buffer || count < 0) return SDSO_E_INVALID;
  if (ctx->peer && ctx->peer->connected && ctx->peer->enabled) {
    PeerExchange* p = ctx->peer;
    if ((size_t)count > p->slot_doubles) return fail(ctx, SDSO_E_INVALID, "sdso_allreduce_f64: count exceeds the exchange block (sdso_peer_alloc)");
    PeerPtrs pp;
    for (int r = 0; r < 16; r++) pp.base[r] = p->base[r];
    p->epoch++;
    if ((p->epoch & 0xffffffffull) == 0) p->epoch++;   // the wire flag 0 means "never written"
    int blocks = (count + 511) / 512;
    if (blocks > 32) blocks = 32;
    if (blocks < 1) blocks = 1;
    peer_allreduce_kernel<<<blocks, 256, 0, ctx->stream>>>(static_cast<double*>(device_buffer), count, pp, p->rank, p->world, (unsigned)(p->epoch & 0xffffffffull), p->slot_doubles, p->d_arrive);
    SDSO_CHECK_LAUNCH(ctx);
    return SDSO_OK;
  }
  if (!ctx->nccl_comm) return fail(ctx, SDSO_E_STATE, "sdso_nccl_init has not been called");
  NcclApi* api = nccl_api(nullptr);
  ncclResult_t r = api->AllReduce(device_buffer, device_buffer, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream);
  if (r != ncclSuccess) return fail(ctx, SDSO_E_CUDA, std::string("ncclAllReduce: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"));
  ctx->launches++;
  return SDSO_OK;
}

}  // extern "C"
