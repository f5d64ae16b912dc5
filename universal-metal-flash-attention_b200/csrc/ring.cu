// ring.cu -- native ring attention (context parallelism over the GPUs of one node): the mfa_ring_* symbols of
// include/mfa_ffi_ext.h.  No reference counterpart ("Reference support: none"); this is BASELINE.json's
// north_star item (4): "for long causal contexts, ring attention that passes K/V blocks over NVLink with NCCL send/recv
// overlapped with compute".
//
// One process per GPU.  The sequence is cut into 2*world chunks, rank r owns chunks r and 2*world-1-r (zig-zag, so causal work
// is balanced) stored next to each other: q / k / v are [B, H, 2C, D] with the low chunk first.  Q stays put; at step s the
// rank works on the K/V pair of rank (r - s) mod world.  With that layout the visible chunk pairs of a step are ONE rectangular
// problem (step_plan), i.e. one attention launch per step through the library's own C entry points:
//     step 0           mfa_attention_forward_ex            causal 2C x 2C on local indices
//     source < rank    mfa_attention_forward_accumulate    2C x C  (both local query chunks see the visitor's low chunk)
//     source > rank    mfa_attention_forward_accumulate     C x 2C (the high query chunk sees both chunks of the visitor)
// the partial (O, L) of a step being merged into the running result inside the attention epilogue.  The backward
// (mfa_ring_attention_backward, end of this file) visits the same rectangles with the dK/dV + dQ kernels.
//
// B200-first shape of the exchange: a B200 has memory to spare (180 GB), so every visiting pair lands in its own slot (no
// double-buffer hand-shake inside a forward), and NVSwitch gives every pair of GPUs full bandwidth, so step s is a DIRECT
// exchange with ranks r +- s instead of a store-and-forward chain: all world-1 transfers of a forward only read the rank's own
// K/V and are queued at once on a highest-priority side stream; step s of the computation waits for slot s alone.
// Two transports:
//   nccl   ncclSend / ncclRecv grouped per step (the north_star's wording).  NCCL is resolved at run time (dlopen of
//          libnccl.so.2: the copy the host process already loaded -- e.g. PyTorch's -- or the system one), so libMFAFFI.so keeps
//          loading on machines without it.  Its kernels are SM resident: they take SMs away from the attention grid while a
//          hop runs.
//   p2p    copy engines: the rank pushes its pair straight into the receiver's slot (peer memory mapped through CUDA IPC,
//          cudaMemcpyAsync over NVLink) and raises the receiver's arrival flag behind it; compute streams wait on the flags
//          with stream memory operations (cuStreamWaitValue32).  No SM and no host thread is involved in a hop.
// (A single persistent launch that walks all slots behind arrival flags was built and measured too -- profiles/r02/
// r02_forward_v2_postmortem.txt: every first-wave work item has to wait for the LAST slot, so it hides far less of the exchange.)
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/mfa_ffi_ext.h"

namespace {

// ---- the slice of the NCCL API the ring needs (types restated so no NCCL header is required at build time) ----------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                 // ncclSuccess = 0
typedef int ncclDataType_t;               // ncclInt8 = 0
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    if (const char* e = getenv("MFA_NCCL_LIBRARY")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      if (h) break;
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) return;
    auto sym = [&](const char* n) { return dlsym(h, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
  });
  return api;
}

// ---- stream memory operations (driver API through the runtime's entry-point query: no libcuda at link time) -----------
typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
struct MemOps { StreamValue32Fn write = nullptr, wait = nullptr; };
MemOps& memops() {
  static MemOps m;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [](const char* name) -> StreamValue32Fn {
      void* ptr = nullptr;
      cudaDriverEntryPointQueryResult qr;
      if (cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
        return reinterpret_cast<StreamValue32Fn>(ptr);
      cudaGetLastError();
      return nullptr;
    };
    m.write = get("cuStreamWriteValue32");
    m.wait = get("cuStreamWaitValue32");
  });
  return m;
}

enum Transport { kNccl = 0, kP2P = 1 };

// what a rank publishes to its peers for the p2p transport
struct PeerBlob {
  cudaIpcMemHandle_t k_visit, v_visit, flags;
  unsigned long long slot_bytes;          // bytes of one K (or V) slot the handles were made for
};

struct Peer {
  char* k_visit = nullptr;
  char* v_visit = nullptr;
  unsigned int* flags = nullptr;
  bool open = false;
};

struct Ring {
  mfa_context_t ctx = nullptr;
  int rank = 0, world = 1, device = 0;
  int transport = kNccl;
  ncclComm_t comm = nullptr;
  bool own_comm = false;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  std::vector<cudaEvent_t> ev_arrived;      // nccl transport: slot s has landed (recorded on the side stream)
  bool have_done = false;
  char* k_visit = nullptr;                  // [world-1] slots of slot_bytes each; slot s at index s-1
  char* v_visit = nullptr;
  size_t slot_bytes = 0;
  // flags[0 .. world)   arrival: slot s holds the pair of forward `epoch`     (written by the sender / the side stream)
  // flags[world .. 2w)  consumed: the pair this rank pushed at step s has been read by its receiver in forward `epoch`
  // flags[2w], [2w+1]   scratch words: sources of the 4-byte peer copies that raise remote flags (side / compute stream)
  unsigned int* flags = nullptr;
  std::vector<Peer> peers;
  bool peers_ready = false;
  unsigned int epoch = 0;
  unsigned long long launches = 0;
  // backward (mfa_ring_attention_backward): its own visiting K/V slots + fp32 work space, NCCL transport
  char* bwd_ws = nullptr;
  size_t bwd_ws_bytes = 0;
  cudaEvent_t ev_part[2] = {nullptr, nullptr}, ev_sent[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> ev_kv_bwd, ev_grad;          // visiting pair s has landed / the gradients of slot s have landed
  std::mutex mu;
};

// dst[bh][row0 + r][:] += src[bh][r][:]  (fp32, D % 4 == 0): gradient partials of a row window into the running gradient
__global__ void add_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n4, long long rows_d4,
                                long long dst_bh_d4, long long row0_d4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long bh = i / rows_d4, j = i - bh * rows_d4;
    float4* d = reinterpret_cast<float4*>(dst) + bh * dst_bh_d4 + row0_d4 + j;
    const float4 a = *d, b = reinterpret_cast<const float4*>(src)[i];
    *d = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}
void add_rows(float* dst, const float* src, long long BH, long long T, long long row0, long long rows, long long D, cudaStream_t st) {
  const long long n4 = BH * rows * D / 4;
  if (n4 <= 0) return;
  long long g = (n4 + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  add_rows_kernel<<<(unsigned)g, 256, 0, st>>>(dst, src, n4, rows * D / 4, T * D / 4, row0 * D / 4);
}

bool debug_on() { const char* d = getenv("MFA_DEBUG"); return d && d[0] && d[0] != '0'; }
#define RDBG(...) do { if (debug_on()) { fprintf(stderr, "[mfa ring] " __VA_ARGS__); fputc('\n', stderr); } } while (0)

// the rectangular problem of ring step `step` on rank `rank` (units of chunks)
struct Plan { int q0, qn, k0, kn; bool causal; };
Plan step_plan(int rank, int world, int step) {
  const int src = ((rank - step) % world + world) % world;
  if (src == rank) return {0, 2, 0, 2, true};
  if (src < rank) return {0, 2, 0, 1, false};
  return {1, 1, 0, 2, false};
}

struct Handle {      // a transient strided device view as an mfa_buffer_t
  mfa_buffer_t h = nullptr;
  Handle(mfa_context_t ctx, void* base, size_t elem_off, size_t esz, int64_t B, int64_t H, int64_t rows, int64_t D, int64_t row_total) {
    const int64_t shape[4] = {B, H, rows, D};
    const int64_t strides[4] = {H * row_total * D, row_total * D, D, 1};
    const size_t span = (size_t)((B - 1) * strides[0] + (H - 1) * strides[1] + (rows - 1) * D + D) * esz;
    mfa_buffer_from_mtl_buffer_with_strides(ctx, reinterpret_cast<char*>(base) + elem_off * esz, span, shape, strides, 4, &h);
  }
  ~Handle() { if (h) mfa_destroy_buffer(h); }
};

struct DeviceScope {
  int prev = -1;
  explicit DeviceScope(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

void close_peers(Ring* r) {
  for (Peer& p : r->peers) {
    if (!p.open) continue;
    if (p.k_visit) cudaIpcCloseMemHandle(p.k_visit);
    if (p.v_visit) cudaIpcCloseMemHandle(p.v_visit);
    if (p.flags) cudaIpcCloseMemHandle(p.flags);
    p = Peer();
  }
  r->peers_ready = false;
  cudaGetLastError();
}

mfa_error_t ring_finish_create(Ring* r, mfa_ring_t* out) {
  cudaGetDevice(&r->device);
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);                  // hi = numerically lowest = highest priority
  bool ok = cudaStreamCreateWithPriority(&r->comm_stream, cudaStreamNonBlocking, hi) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&r->ev_ready, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&r->ev_done, cudaEventDisableTiming) == cudaSuccess;
  r->ev_arrived.assign((size_t)r->world, nullptr);
  for (int s = 1; s < r->world && ok; ++s) ok = cudaEventCreateWithFlags(&r->ev_arrived[s], cudaEventDisableTiming) == cudaSuccess;
  const size_t nflags = 2 * (size_t)r->world + 2;
  ok = ok && cudaMalloc(&r->flags, sizeof(unsigned int) * nflags) == cudaSuccess;
  ok = ok && cudaMemset(r->flags, 0, sizeof(unsigned int) * nflags) == cudaSuccess;
  r->peers.assign((size_t)r->world, Peer());
  if (!ok) { cudaGetLastError(); mfa_ring_destroy(reinterpret_cast<mfa_ring_t>(r)); return MFA_ERROR_EXECUTION_FAILED; }
  if (const char* e = getenv("MFA_RING_TRANSPORT")) r->transport = strcmp(e, "p2p") == 0 ? kP2P : kNccl;
  *out = reinterpret_cast<mfa_ring_t>(r);
  return MFA_SUCCESS;
}

mfa_error_t ensure_slots(Ring* r, size_t slot_bytes) {
  if (r->world <= 1 || r->slot_bytes >= slot_bytes) return MFA_SUCCESS;
  if (r->transport == kP2P && r->peers_ready) return MFA_ERROR_INVALID_ARGS;   // peers hold handles of the old slots: mfa_ring_prepare first
  if (r->have_done) cudaEventSynchronize(r->ev_done);
  if (r->k_visit) cudaFree(r->k_visit);
  if (r->v_visit) cudaFree(r->v_visit);
  r->k_visit = r->v_visit = nullptr; r->slot_bytes = 0;
  const size_t total = slot_bytes * (size_t)(r->world - 1);
  if (cudaMalloc(&r->k_visit, total) != cudaSuccess || cudaMalloc(&r->v_visit, total) != cudaSuccess) { cudaGetLastError(); return MFA_ERROR_MEMORY_ALLOCATION; }
  r->slot_bytes = slot_bytes;
  return MFA_SUCCESS;
}

}  // namespace

extern "C" {

bool mfa_ring_transport_available(void) { return nccl().ok; }

mfa_error_t mfa_ring_get_unique_id(void* id_out, size_t id_bytes) {
  if (!id_out || id_bytes < sizeof(ncclUniqueId)) return MFA_ERROR_INVALID_ARGS;
  if (!nccl().ok) return MFA_ERROR_DEVICE_NOT_SUPPORTED;
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != 0) return MFA_ERROR_EXECUTION_FAILED;
  memcpy(id_out, &id, sizeof(id));
  return MFA_SUCCESS;
}

mfa_error_t mfa_ring_create(mfa_context_t context, const void* unique_id, size_t id_bytes, int32_t rank, int32_t world_size,
                            mfa_ring_t* ring) {
  if (!context || !ring || rank < 0 || world_size < 1 || rank >= world_size || world_size > 255) return MFA_ERROR_INVALID_ARGS;
  *ring = nullptr;
  Ring* r = new (std::nothrow) Ring();
  if (!r) return MFA_ERROR_MEMORY_ALLOCATION;
  r->ctx = context; r->rank = rank; r->world = world_size;
  if (world_size > 1 && unique_id) {                 // a NULL id creates a ring for the p2p transport only
    if (id_bytes < sizeof(ncclUniqueId)) { delete r; return MFA_ERROR_INVALID_ARGS; }
    if (!nccl().ok) { delete r; return MFA_ERROR_DEVICE_NOT_SUPPORTED; }
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    const ncclResult_t rc = nccl().CommInitRank(&r->comm, world_size, id, rank);
    if (rc != 0) {
      RDBG("ncclCommInitRank: %s", nccl().GetErrorString ? nccl().GetErrorString(rc) : "error");
      delete r;
      return MFA_ERROR_EXECUTION_FAILED;
    }
    r->own_comm = true;
  }
  mfa_error_t e = ring_finish_create(r, ring);
  if (e == MFA_SUCCESS && world_size > 1 && !unique_id) r->transport = kP2P;
  return e;
}

mfa_error_t mfa_ring_create_from_comm(mfa_context_t context, void* nccl_comm, int32_t rank, int32_t world_size, mfa_ring_t* ring) {
  if (!context || !ring || rank < 0 || world_size < 1 || rank >= world_size || world_size > 255) return MFA_ERROR_INVALID_ARGS;
  *ring = nullptr;
  if (world_size > 1 && !nccl_comm) return MFA_ERROR_INVALID_ARGS;
  if (world_size > 1 && !nccl().ok) return MFA_ERROR_DEVICE_NOT_SUPPORTED;
  Ring* r = new (std::nothrow) Ring();
  if (!r) return MFA_ERROR_MEMORY_ALLOCATION;
  r->ctx = context; r->rank = rank; r->world = world_size;
  r->comm = reinterpret_cast<ncclComm_t>(nccl_comm);
  return ring_finish_create(r, ring);
}

void mfa_ring_destroy(mfa_ring_t ring) {
  if (!ring) return;
  Ring* r = reinterpret_cast<Ring*>(ring);
  DeviceScope ds(r->device);
  if (r->comm_stream) cudaStreamSynchronize(r->comm_stream);
  close_peers(r);
  if (r->own_comm && r->comm) nccl().CommDestroy(r->comm);
  if (r->k_visit) cudaFree(r->k_visit);
  if (r->v_visit) cudaFree(r->v_visit);
  if (r->flags) cudaFree(r->flags);
  if (r->bwd_ws) cudaFree(r->bwd_ws);
  for (cudaEvent_t e : r->ev_part) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : r->ev_sent) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : r->ev_kv_bwd) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : r->ev_grad) if (e) cudaEventDestroy(e);
  if (r->ev_ready) cudaEventDestroy(r->ev_ready);
  if (r->ev_done) cudaEventDestroy(r->ev_done);
  for (cudaEvent_t e : r->ev_arrived) if (e) cudaEventDestroy(e);
  if (r->comm_stream) cudaStreamDestroy(r->comm_stream);
  cudaGetLastError();
  delete r;
}

uint64_t mfa_ring_launch_count(mfa_ring_t ring) { return ring ? reinterpret_cast<Ring*>(ring)->launches : 0; }
int32_t mfa_ring_transport(mfa_ring_t ring) { return ring ? reinterpret_cast<Ring*>(ring)->transport : -1; }

// ---- p2p transport set-up: allocate the slots for the largest problem to come, publish their IPC handles, open the peers'.
// prepare -> export -> (the host gathers every rank's blob, any channel) -> import; all three are collective in effect.
size_t mfa_ring_handle_bytes(void) { return sizeof(PeerBlob); }

mfa_error_t mfa_ring_prepare(mfa_ring_t ring, uint32_t batch_size, uint32_t chunk_rows, uint32_t num_heads, uint16_t head_dim) {
  if (!ring) return MFA_ERROR_INVALID_ARGS;
  Ring* r = reinterpret_cast<Ring*>(ring);
  std::lock_guard<std::mutex> lock(r->mu);
  DeviceScope ds(r->device);
  close_peers(r);
  return ensure_slots(r, (size_t)batch_size * num_heads * 2 * chunk_rows * head_dim * 2);
}

mfa_error_t mfa_ring_export_handles(mfa_ring_t ring, void* blob_out, size_t blob_bytes) {
  if (!ring || !blob_out || blob_bytes < sizeof(PeerBlob)) return MFA_ERROR_INVALID_ARGS;
  Ring* r = reinterpret_cast<Ring*>(ring);
  std::lock_guard<std::mutex> lock(r->mu);
  if (r->world > 1 && (!r->k_visit || !r->v_visit)) return MFA_ERROR_INVALID_ARGS;
  DeviceScope ds(r->device);
  PeerBlob b;
  memset(&b, 0, sizeof(b));
  b.slot_bytes = r->slot_bytes;
  if (r->world > 1 && (cudaIpcGetMemHandle(&b.k_visit, r->k_visit) != cudaSuccess || cudaIpcGetMemHandle(&b.v_visit, r->v_visit) != cudaSuccess ||
                       cudaIpcGetMemHandle(&b.flags, r->flags) != cudaSuccess)) {
    RDBG("cudaIpcGetMemHandle: %s", cudaGetErrorString(cudaGetLastError()));
    return MFA_ERROR_EXECUTION_FAILED;
  }
  memcpy(blob_out, &b, sizeof(b));
  return MFA_SUCCESS;
}

mfa_error_t mfa_ring_import_handles(mfa_ring_t ring, const void* blobs, size_t bytes_per_rank) {
  if (!ring || !blobs || bytes_per_rank < sizeof(PeerBlob)) return MFA_ERROR_INVALID_ARGS;
  Ring* r = reinterpret_cast<Ring*>(ring);
  std::lock_guard<std::mutex> lock(r->mu);
  DeviceScope ds(r->device);
  close_peers(r);
  if (!memops().write || !memops().wait) return MFA_ERROR_DEVICE_NOT_SUPPORTED;
  for (int p = 0; p < r->world; ++p) {
    if (p == r->rank) continue;
    PeerBlob b;
    memcpy(&b, reinterpret_cast<const char*>(blobs) + (size_t)p * bytes_per_rank, sizeof(b));
    if (b.slot_bytes != r->slot_bytes) return MFA_ERROR_INVALID_ARGS;           // every rank prepares for the same problem
    Peer& pe = r->peers[p];
    void *a = nullptr, *c = nullptr, *f = nullptr;
    if (cudaIpcOpenMemHandle(&a, b.k_visit, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&c, b.v_visit, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&f, b.flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      RDBG("cudaIpcOpenMemHandle (peer %d): %s", p, cudaGetErrorString(cudaGetLastError()));
      close_peers(r);
      return MFA_ERROR_EXECUTION_FAILED;
    }
    pe.k_visit = reinterpret_cast<char*>(a); pe.v_visit = reinterpret_cast<char*>(c); pe.flags = reinterpret_cast<unsigned int*>(f);
    pe.open = true;
  }
  r->peers_ready = true;
  r->transport = kP2P;
  return MFA_SUCCESS;
}

// q, k, v: this rank's [low | high] chunk pair, device-resident, contiguous [B, H, 2C, D] in `precision` (bf16 / fp16).
// out: fp32 [B, H, 2C, D], lse: fp32 [B, H, 2C] (log2 units).  Attention launches go to `stream` (the exchange to the ring's own
// side stream); the call returns without synchronising unless stream is NULL.
mfa_error_t mfa_ring_attention_forward(mfa_ring_t ring, mfa_buffer_t q, mfa_buffer_t k, mfa_buffer_t v, mfa_buffer_t out,
                                       mfa_buffer_t lse, uint32_t batch_size, uint32_t chunk_rows, uint32_t num_heads,
                                       uint16_t head_dim, float softmax_scale, mfa_precision_t precision, void* stream) {
  if (!ring || !q || !k || !v || !out || !lse) return MFA_ERROR_INVALID_ARGS;
  if (precision != MFA_PRECISION_BF16 && precision != MFA_PRECISION_FP16) return MFA_ERROR_INVALID_ARGS;
  Ring* r = reinterpret_cast<Ring*>(ring);
  std::lock_guard<std::mutex> lock(r->mu);
  const int64_t B = batch_size, H = num_heads, C = chunk_rows, T = 2 * C, D = head_dim;
  const size_t n = (size_t)B * H * T * D, esz = 2, bytes = n * esz;
  if (n == 0) return MFA_SUCCESS;
  void* qd = mfa_buffer_contents(q);
  char* kd = reinterpret_cast<char*>(mfa_buffer_contents(k));
  char* vd = reinterpret_cast<char*>(mfa_buffer_contents(v));
  if (!qd || !kd || !vd) return MFA_ERROR_INVALID_ARGS;
  DeviceScope ds(r->device);
  const int W = r->world;
  if (W > 1) {
    if (r->transport == kNccl && !r->comm) return MFA_ERROR_INVALID_ARGS;
    if (r->transport == kP2P && (!r->peers_ready || r->slot_bytes < bytes)) return MFA_ERROR_INVALID_ARGS;
    if (mfa_error_t e = ensure_slots(r, bytes); e != MFA_SUCCESS) return e;
  }
  cudaStream_t cs = reinterpret_cast<cudaStream_t>(stream);
  const bool blocking = stream == nullptr;
  cudaStream_t own = nullptr;
  if (blocking) { if (cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking) != cudaSuccess) return MFA_ERROR_EXECUTION_FAILED; cs = own; }
  mfa_error_t rc = MFA_SUCCESS;
  const unsigned int epoch = ++r->epoch;
  unsigned int* arrived = r->flags;
  unsigned int* consumed = r->flags + W;
  unsigned int* scratch = r->flags + 2 * W;

  // ---- the exchange: every transfer of this forward reads the rank's own pair only, so all of them are queued now
  if (W > 1) {
    cudaEventRecord(r->ev_ready, cs);                          // the caller's K / V were produced on the compute stream
    cudaStreamWaitEvent(r->comm_stream, r->ev_ready, 0);
    if (r->transport == kNccl) {
      if (r->have_done) cudaStreamWaitEvent(r->comm_stream, r->ev_done, 0);     // the previous forward still reads the slots
      NcclApi& nc = nccl();
      for (int s = 1; s < W && rc == MFA_SUCCESS; ++s) {
        const int dst = (r->rank + s) % W, src = (r->rank - s + W) % W;
        ncclResult_t e = nc.GroupStart();
        if (e == 0) e = nc.Send(kd, bytes, 0, dst, r->comm, r->comm_stream);
        if (e == 0) e = nc.Send(vd, bytes, 0, dst, r->comm, r->comm_stream);
        if (e == 0) e = nc.Recv(r->k_visit + (size_t)(s - 1) * r->slot_bytes, bytes, 0, src, r->comm, r->comm_stream);
        if (e == 0) e = nc.Recv(r->v_visit + (size_t)(s - 1) * r->slot_bytes, bytes, 0, src, r->comm, r->comm_stream);
        const ncclResult_t e2 = nc.GroupEnd();
        if (e != 0 || e2 != 0) { RDBG("nccl send/recv: %s", nc.GetErrorString ? nc.GetErrorString(e ? e : e2) : "error"); rc = MFA_ERROR_EXECUTION_FAILED; }
        cudaEventRecord(r->ev_arrived[s], r->comm_stream);
      }
    } else {
      MemOps& mo = memops();
      mo.write(r->comm_stream, reinterpret_cast<CUdeviceptr>(scratch), epoch, 0);        // source word of the remote flag writes
      for (int s = 1; s < W && rc == MFA_SUCCESS; ++s) {
        const int dst = (r->rank + s) % W;
        Peer& pe = r->peers[dst];
        // the receiver has read what this rank pushed into its slot s in the previous forward
        if (epoch > 1) mo.wait(r->comm_stream, reinterpret_cast<CUdeviceptr>(consumed + s), epoch - 1, CU_STREAM_WAIT_VALUE_GEQ);
        cudaError_t e = cudaMemcpyAsync(pe.k_visit + (size_t)(s - 1) * r->slot_bytes, kd, bytes, cudaMemcpyDefault, r->comm_stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(pe.v_visit + (size_t)(s - 1) * r->slot_bytes, vd, bytes, cudaMemcpyDefault, r->comm_stream);
        // ... and only then its arrival flag (stream order): a 4-byte peer copy of the epoch word
        if (e == cudaSuccess) e = cudaMemcpyAsync(pe.flags + s, scratch, sizeof(unsigned int), cudaMemcpyDefault, r->comm_stream);
        if (e != cudaSuccess) { RDBG("peer copy: %s", cudaGetErrorString(e)); cudaGetLastError(); rc = MFA_ERROR_EXECUTION_FAILED; }
      }
    }
  }

  // ---- the computation: one launch per step on the compute stream
  char* cur_k = kd;
  char* cur_v = vd;
  for (int step = 0; step < W && rc == MFA_SUCCESS; ++step) {
    if (step > 0) {
      if (r->transport == kNccl) cudaStreamWaitEvent(cs, r->ev_arrived[step], 0);
      else memops().wait(cs, reinterpret_cast<CUdeviceptr>(arrived + step), epoch, CU_STREAM_WAIT_VALUE_GEQ);
      cur_k = r->k_visit + (size_t)(step - 1) * r->slot_bytes;
      cur_v = r->v_visit + (size_t)(step - 1) * r->slot_bytes;
    }
    const Plan pl = step_plan(r->rank, W, step);
    {
      Handle hq(r->ctx, qd, (size_t)pl.q0 * C * D, esz, B, H, (int64_t)pl.qn * C, D, T);
      Handle hk(r->ctx, cur_k, (size_t)pl.k0 * C * D, esz, B, H, (int64_t)pl.kn * C, D, T);
      Handle hv(r->ctx, cur_v, (size_t)pl.k0 * C * D, esz, B, H, (int64_t)pl.kn * C, D, T);
      if (!hq.h || !hk.h || !hv.h) { rc = MFA_ERROR_MEMORY_ALLOCATION; break; }
      if (step == 0)
        rc = mfa_attention_forward_ex(r->ctx, hq.h, hk.h, hv.h, out, lse, batch_size, (uint32_t)(pl.qn * C), (uint32_t)(pl.kn * C),
                                      num_heads, head_dim, softmax_scale, pl.causal, -1, precision, MFA_PRECISION_FP32,
                                      nullptr, 0, nullptr, nullptr, 0, MFA_MASK_TYPE_NONE, MFA_MASK_SCALAR_BYTE, cs);
      else
        rc = mfa_attention_forward_accumulate(r->ctx, hq.h, hk.h, hv.h, out, lse, batch_size, (uint32_t)(pl.qn * C),
                                              (uint32_t)(pl.kn * C), num_heads, head_dim, softmax_scale, pl.causal, -1, precision,
                                              (uint32_t)(pl.q0 * C), (uint32_t)T, cs);
      ++r->launches;
    }
    if (rc != MFA_SUCCESS) break;
    if (step > 0 && r->transport == kP2P) {
      // tell the sender of this slot (rank - step) that it may be overwritten by the next forward: its consumed[step]
      const int src = (r->rank - step + W) % W;
      memops().write(cs, reinterpret_cast<CUdeviceptr>(scratch + 1), epoch, 0);      // the compute stream's own source word
      cudaError_t e = cudaMemcpyAsync(r->peers[src].flags + W + step, scratch + 1, sizeof(unsigned int), cudaMemcpyDefault, cs);
      if (e != cudaSuccess) { cudaGetLastError(); rc = MFA_ERROR_EXECUTION_FAILED; }
    }
  }
  if (rc == MFA_SUCCESS && W > 1) { cudaEventRecord(r->ev_done, cs); r->have_done = true; }
  if (blocking) {
    if (cudaStreamSynchronize(cs) != cudaSuccess) { cudaGetLastError(); rc = rc == MFA_SUCCESS ? MFA_ERROR_EXECUTION_FAILED : rc; }
    cudaStreamDestroy(own);
  }
  return rc;
}

// Backward of mfa_ring_attention_forward.  q, k, v as in the forward; out / lse are the forward's results of this
// rank; dout is the upstream gradient in `precision`, contiguous [B, H, 2C, D]; dq, dk, dv fp32 [B, H, 2C, D] (overwritten).
// The same rectangles as the forward: with the final L and D = scale * rowsum(dO * O) of a query row the flash backward of a
// (query block, key block) pair is an exact partial sum, so step s runs the dK/dV + dQ kernels on (local rows, visiting pair),
// adds its dQ into the local gradient and sends the pair's dK / dV DIRECTLY back to its owner (rank - s), which adds them into
// its own -- the mirror image of the forward's direct exchange.  NCCL transport (the ring needs a communicator); the K / V
// exchange and the gradient returns run on the side stream behind events, the backward kernels on `stream`.
mfa_error_t mfa_ring_attention_backward(mfa_ring_t ring, mfa_buffer_t q, mfa_buffer_t k, mfa_buffer_t v, mfa_buffer_t out,
                                        mfa_buffer_t lse, mfa_buffer_t dout, mfa_buffer_t dq, mfa_buffer_t dk, mfa_buffer_t dv,
                                        uint32_t batch_size, uint32_t chunk_rows, uint32_t num_heads, uint16_t head_dim,
                                        float softmax_scale, mfa_precision_t precision, void* stream) {
  if (!ring || !q || !k || !v || !out || !lse || !dout || !dq || !dk || !dv) return MFA_ERROR_INVALID_ARGS;
  if (precision != MFA_PRECISION_BF16 && precision != MFA_PRECISION_FP16) return MFA_ERROR_INVALID_ARGS;
  Ring* r = reinterpret_cast<Ring*>(ring);
  std::lock_guard<std::mutex> lock(r->mu);
  const int64_t B = batch_size, H = num_heads, C = chunk_rows, T = 2 * C, D = head_dim, BH = B * H;
  const size_t n = (size_t)BH * T * D, esz = 2, bytes = n * esz, gbytes = n * 4, half16 = n / 2 * esz, half32 = n / 2 * 4;
  if (n == 0) return MFA_SUCCESS;
  if (D % 4) return MFA_ERROR_INVALID_ARGS;
  char* qd = reinterpret_cast<char*>(mfa_buffer_contents(q));
  char* kd = reinterpret_cast<char*>(mfa_buffer_contents(k));
  char* vd = reinterpret_cast<char*>(mfa_buffer_contents(v));
  char* od = reinterpret_cast<char*>(mfa_buffer_contents(out));
  char* ld = reinterpret_cast<char*>(mfa_buffer_contents(lse));
  char* gd = reinterpret_cast<char*>(mfa_buffer_contents(dout));
  float* dqd = reinterpret_cast<float*>(mfa_buffer_contents(dq));
  float* dkd = reinterpret_cast<float*>(mfa_buffer_contents(dk));
  float* dvd = reinterpret_cast<float*>(mfa_buffer_contents(dv));
  if (!qd || !kd || !vd || !od || !ld || !gd || !dqd || !dkd || !dvd) return MFA_ERROR_INVALID_ARGS;
  DeviceScope ds(r->device);
  const int W = r->world;
  if (W > 1 && (!r->comm || !nccl().ok)) return MFA_ERROR_DEVICE_NOT_SUPPORTED;
  cudaStream_t cs = reinterpret_cast<cudaStream_t>(stream);
  const bool blocking = stream == nullptr;
  cudaStream_t own = nullptr;
  if (blocking) { if (cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking) != cudaSuccess) return MFA_ERROR_EXECUTION_FAILED; cs = own; }
  mfa_error_t rc = MFA_SUCCESS;
  auto done = [&](mfa_error_t e) {
    if (blocking) { if (cudaStreamSynchronize(cs) != cudaSuccess) { cudaGetLastError(); if (e == MFA_SUCCESS) e = MFA_ERROR_EXECUTION_FAILED; } cudaStreamDestroy(own); }
    return e;
  };

  // step 0: the rank's own pair, causal on local indices, straight into dq / dk / dv
  rc = mfa_attention_backward_ex(r->ctx, dout, q, k, v, out, lse, dq, dk, dv, nullptr, batch_size, (uint32_t)T, (uint32_t)T, num_heads,
                                 head_dim, softmax_scale, true, -1, precision, nullptr, 0, nullptr, nullptr, 0, MFA_MASK_TYPE_NONE,
                                 MFA_MASK_SCALAR_BYTE, cs);
  ++r->launches;
  if (W == 1 || rc != MFA_SUCCESS) return done(rc);

  // ---- work space: visiting K / V slots, the gathered high-chunk windows, gradient partials (two send sets, W - 1 receive sets)
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_kv = 0, o_win = o_kv + 2 * (size_t)(W - 1) * up(bytes);
  const size_t o_qh = o_win, o_gh = o_qh + up(half16), o_oh = o_gh + up(half16), o_lh = o_oh + up(half32);
  const size_t o_dqp = o_lh + up((size_t)BH * C * 4), o_send = o_dqp + up(gbytes), o_recv = o_send + 4 * up(gbytes);
  const size_t total = o_recv + 2 * (size_t)(W - 1) * up(gbytes);
  if (r->bwd_ws_bytes < total) {
    cudaDeviceSynchronize();
    if (r->bwd_ws) cudaFree(r->bwd_ws);
    r->bwd_ws = nullptr; r->bwd_ws_bytes = 0;
    if (cudaMalloc(&r->bwd_ws, total) != cudaSuccess) { cudaGetLastError(); return done(MFA_ERROR_MEMORY_ALLOCATION); }
    r->bwd_ws_bytes = total;
  }
  if (r->ev_kv_bwd.empty()) {
    bool ok = true;
    r->ev_kv_bwd.assign((size_t)W, nullptr); r->ev_grad.assign((size_t)W, nullptr);
    for (int s = 1; s < W && ok; ++s)
      ok = cudaEventCreateWithFlags(&r->ev_kv_bwd[s], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&r->ev_grad[s], cudaEventDisableTiming) == cudaSuccess;
    for (int j = 0; j < 2 && ok; ++j)
      ok = cudaEventCreateWithFlags(&r->ev_part[j], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&r->ev_sent[j], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { cudaGetLastError(); return done(MFA_ERROR_EXECUTION_FAILED); }
  }
  char* ws = r->bwd_ws;
  auto kslot = [&](int s) { return ws + o_kv + (size_t)(2 * (s - 1)) * up(bytes); };
  auto vslot = [&](int s) { return ws + o_kv + (size_t)(2 * (s - 1) + 1) * up(bytes); };
  auto send_dk = [&](int j) { return reinterpret_cast<float*>(ws + o_send + (size_t)(2 * j) * up(gbytes)); };
  auto send_dv = [&](int j) { return reinterpret_cast<float*>(ws + o_send + (size_t)(2 * j + 1) * up(gbytes)); };
  auto recv_dk = [&](int s) { return reinterpret_cast<float*>(ws + o_recv + (size_t)(2 * (s - 1)) * up(gbytes)); };
  auto recv_dv = [&](int s) { return reinterpret_cast<float*>(ws + o_recv + (size_t)(2 * (s - 1) + 1) * up(gbytes)); };
  float* dq_part = reinterpret_cast<float*>(ws + o_dqp);
  NcclApi& nc = nccl();

  // ---- K / V exchange: all hops queued now on the side stream (they only read the rank's own pair)
  cudaEventRecord(r->ev_ready, cs);
  cudaStreamWaitEvent(r->comm_stream, r->ev_ready, 0);
  for (int s = 1; s < W && rc == MFA_SUCCESS; ++s) {
    const int dst = (r->rank + s) % W, src = (r->rank - s + W) % W;
    ncclResult_t e = nc.GroupStart();
    if (e == 0) e = nc.Send(kd, bytes, 0, dst, r->comm, r->comm_stream);
    if (e == 0) e = nc.Send(vd, bytes, 0, dst, r->comm, r->comm_stream);
    if (e == 0) e = nc.Recv(kslot(s), bytes, 0, src, r->comm, r->comm_stream);
    if (e == 0) e = nc.Recv(vslot(s), bytes, 0, src, r->comm, r->comm_stream);
    const ncclResult_t e2 = nc.GroupEnd();
    if (e != 0 || e2 != 0) rc = MFA_ERROR_EXECUTION_FAILED;
    cudaEventRecord(r->ev_kv_bwd[s], r->comm_stream);
  }
  if (rc != MFA_SUCCESS) return done(rc);

  // ---- the high query chunk as contiguous operands (rows C .. 2C of every (b, h)): used by every step whose source rank > rank
  const size_t rq = (size_t)T * D * esz, ro = (size_t)T * D * 4, rl = (size_t)T * 4;
  cudaMemcpy2DAsync(ws + o_qh, rq / 2, qd + rq / 2, rq, rq / 2, (size_t)BH, cudaMemcpyDeviceToDevice, cs);
  cudaMemcpy2DAsync(ws + o_gh, rq / 2, gd + rq / 2, rq, rq / 2, (size_t)BH, cudaMemcpyDeviceToDevice, cs);
  cudaMemcpy2DAsync(ws + o_oh, ro / 2, od + ro / 2, ro, ro / 2, (size_t)BH, cudaMemcpyDeviceToDevice, cs);
  cudaMemcpy2DAsync(ws + o_lh, rl / 2, ld + rl / 2, rl, rl / 2, (size_t)BH, cudaMemcpyDeviceToDevice, cs);
  struct Dev { mfa_buffer_t h = nullptr; Dev(mfa_context_t c, void* p, size_t b) { mfa_buffer_from_mtl_buffer(c, p, b, &h); } ~Dev() { if (h) mfa_destroy_buffer(h); } };
  Dev h_qh(r->ctx, ws + o_qh, half16), h_gh(r->ctx, ws + o_gh, half16), h_oh(r->ctx, ws + o_oh, half32), h_lh(r->ctx, ws + o_lh, (size_t)BH * C * 4);
  Dev h_dqp(r->ctx, dq_part, gbytes);
  if (!h_qh.h || !h_gh.h || !h_oh.h || !h_lh.h || !h_dqp.h) return done(MFA_ERROR_MEMORY_ALLOCATION);

  for (int step = 1; step < W && rc == MFA_SUCCESS; ++step) {
    const int j = step & 1;
    const int src = (r->rank - step + W) % W, from = (r->rank + step) % W;      // owner of the visiting pair / whose gradients come back
    const Plan pl = step_plan(r->rank, W, step);
    cudaStreamWaitEvent(cs, r->ev_kv_bwd[step], 0);
    if (step > 2) cudaStreamWaitEvent(cs, r->ev_sent[j], 0);                   // send set j is free again
    {
      Handle hk(r->ctx, kslot(step), (size_t)pl.k0 * C * D, esz, B, H, (int64_t)pl.kn * C, D, T);
      Handle hv(r->ctx, vslot(step), (size_t)pl.k0 * C * D, esz, B, H, (int64_t)pl.kn * C, D, T);
      Dev h_dk(r->ctx, send_dk(j), gbytes), h_dv(r->ctx, send_dv(j), gbytes);
      if (!hk.h || !hv.h || !h_dk.h || !h_dv.h) { rc = MFA_ERROR_MEMORY_ALLOCATION; break; }
      const bool all_rows = pl.qn == 2;
      rc = mfa_attention_backward_ex(r->ctx, all_rows ? dout : h_gh.h, all_rows ? q : h_qh.h, hk.h, hv.h, all_rows ? out : h_oh.h,
                                     all_rows ? lse : h_lh.h, h_dqp.h, h_dk.h, h_dv.h, nullptr, batch_size, (uint32_t)(pl.qn * C),
                                     (uint32_t)(pl.kn * C), num_heads, head_dim, softmax_scale, false, -1, precision, nullptr, 0,
                                     nullptr, nullptr, 0, MFA_MASK_TYPE_NONE, MFA_MASK_SCALAR_BYTE, cs);
      ++r->launches;
      if (rc != MFA_SUCCESS) break;
    }
    add_rows(dqd, dq_part, BH, T, (long long)pl.q0 * C, (long long)pl.qn * C, D, cs);
    cudaEventRecord(r->ev_part[j], cs);
    // gradients of the visiting pair go home; the pair this rank lent to rank + step comes back with what that rank computed:
    // it saw the low chunk only when it is the higher rank, both chunks otherwise
    const size_t send_bytes = (size_t)BH * pl.kn * C * D * 4;
    const int back_kn = from > r->rank ? 1 : 2;
    const size_t recv_bytes = (size_t)BH * back_kn * C * D * 4;
    cudaStreamWaitEvent(r->comm_stream, r->ev_part[j], 0);
    ncclResult_t e = nc.GroupStart();
    if (e == 0) e = nc.Send(send_dk(j), send_bytes, 0, src, r->comm, r->comm_stream);
    if (e == 0) e = nc.Send(send_dv(j), send_bytes, 0, src, r->comm, r->comm_stream);
    if (e == 0) e = nc.Recv(recv_dk(step), recv_bytes, 0, from, r->comm, r->comm_stream);
    if (e == 0) e = nc.Recv(recv_dv(step), recv_bytes, 0, from, r->comm, r->comm_stream);
    const ncclResult_t e2 = nc.GroupEnd();
    if (e != 0 || e2 != 0) { rc = MFA_ERROR_EXECUTION_FAILED; break; }
    cudaEventRecord(r->ev_sent[j], r->comm_stream);
    cudaEventRecord(r->ev_grad[step], r->comm_stream);
  }
  // ---- what the other ranks computed for this rank's K / V
  for (int step = 1; step < W && rc == MFA_SUCCESS; ++step) {
    const int from = (r->rank + step) % W;
    const int back_kn = from > r->rank ? 1 : 2;
    cudaStreamWaitEvent(cs, r->ev_grad[step], 0);
    add_rows(dkd, recv_dk(step), BH, T, 0, (long long)back_kn * C, D, cs);
    add_rows(dvd, recv_dv(step), BH, T, 0, (long long)back_kn * C, D, cs);
  }
  if (cudaGetLastError() != cudaSuccess && rc == MFA_SUCCESS) rc = MFA_ERROR_EXECUTION_FAILED;
  // the side stream's last sends read the send sets: the next call may not reuse them before (same stream order on comm_stream)
  return done(rc);
}

}  // extern "C"
