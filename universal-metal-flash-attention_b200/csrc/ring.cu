// ring.cu -- native ring attention (context parallelism over the GPUs of one node): the mfa_ring_* symbols of
// include/mfa_ffi_ext.h.  No reference counterpart ("Reference support: none"); this is BASELINE.json's
// north_star item (4): "for long causal contexts, ring attention that passes K/V blocks over NVLink with NCCL send/recv
// overlapped with compute".
//
// One process per GPU.  The sequence is cut into 2*world chunks, rank r owns chunks r and 2*world-1-r (zig-zag, so causal work
// is balanced) stored next to each other: q / k / v are [B, H, 2C, D] with the low chunk first.  Q stays put; the K/V pair of
// rank (r - s) mod world reaches rank r at step s = 1 .. world-1.
//
// B200-first shape of the algorithm: a B200 has memory to spare (180 GB), so the visiting pairs are not double-buffered and
// thrown away -- every step lands in its own slot -- and NVSwitch gives every pair of GPUs full bandwidth, so step s is a direct
// exchange with rank r +- s (ncclSend / ncclRecv grouped per step on a highest-priority side stream), not a store-and-forward
// chain.  The compute side is then ONE persistent attention launch per forward (attn_fwd_tc.cu, launch_fwd_tc_ring): every work
// item walks its own causal block and then the visiting slots in arrival order, the kernel's TMA producer warps polling a
// per-slot arrival flag the side stream raises behind each step.  O stays in TMEM across all ring steps: no partial (O, L) is
// merged or written, there is no per-step launch, and the exchange overlaps the whole computation.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the host process already loaded -- e.g. PyTorch's -- or the
// system one), so libMFAFFI.so keeps loading on machines without NCCL.  NCCL's send / recv kernels are SM resident and the
// attention grid is persistent, so the launch leaves `reserve_sms` SMs (default 8) to the transport while hops are in flight.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

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

// arrival flag of a slot, raised on the transport stream behind the step's receives
__global__ void ring_raise_flag_kernel(unsigned int* flag, unsigned int value) {
  *reinterpret_cast<volatile unsigned int*>(flag) = value;
  __threadfence_system();
}

typedef CUresult (*StreamWriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWriteValue32Fn stream_write_value32() {
  static StreamWriteValue32Fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    if (getenv("MFA_RING_FLAG_KERNEL")) return;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &ptr, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<StreamWriteValue32Fn>(ptr);
    else
      cudaGetLastError();
  });
  return fn;
}

struct Ring {
  mfa_context_t ctx = nullptr;
  int rank = 0, world = 1, device = 0;
  ncclComm_t comm = nullptr;
  bool own_comm = false;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  bool have_done = false;
  void* k_visit = nullptr;                  // [world-1][B][H][2C][D]
  void* v_visit = nullptr;
  size_t visit_cap = 0;                     // bytes per array
  unsigned int* flags = nullptr;            // [world] device words
  unsigned int epoch = 0;
  int reserve_sms = 8;
  unsigned long long launches = 0;
  std::mutex mu;
};

bool debug_on() { const char* d = getenv("MFA_DEBUG"); return d && d[0] && d[0] != '0'; }
#define RDBG(...) do { if (debug_on()) { fprintf(stderr, "[mfa ring] " __VA_ARGS__); fputc('\n', stderr); } } while (0)

mfa_error_t ring_finish_create(Ring* r, mfa_ring_t* out) {
  cudaGetDevice(&r->device);
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);                  // hi = numerically lowest = highest priority
  bool ok = cudaStreamCreateWithPriority(&r->comm_stream, cudaStreamNonBlocking, hi) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&r->ev_ready, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&r->ev_done, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaMalloc(&r->flags, sizeof(unsigned int) * (size_t)(r->world + 1)) == cudaSuccess;
  ok = ok && cudaMemset(r->flags, 0, sizeof(unsigned int) * (size_t)(r->world + 1)) == cudaSuccess;
  if (!ok) { cudaGetLastError(); mfa_ring_destroy(reinterpret_cast<mfa_ring_t>(r)); return MFA_ERROR_EXECUTION_FAILED; }
  if (const char* e = getenv("MFA_RING_RESERVE_SMS")) r->reserve_sms = atoi(e);
  *out = reinterpret_cast<mfa_ring_t>(r);
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
  if (world_size > 1) {
    if (!unique_id || id_bytes < sizeof(ncclUniqueId)) { delete r; return MFA_ERROR_INVALID_ARGS; }
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
  return ring_finish_create(r, ring);
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
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != r->device) cudaSetDevice(r->device);
  if (r->comm_stream) cudaStreamSynchronize(r->comm_stream);
  if (r->own_comm && r->comm) nccl().CommDestroy(r->comm);
  if (r->k_visit) cudaFree(r->k_visit);
  if (r->v_visit) cudaFree(r->v_visit);
  if (r->flags) cudaFree(r->flags);
  if (r->ev_ready) cudaEventDestroy(r->ev_ready);
  if (r->ev_done) cudaEventDestroy(r->ev_done);
  if (r->comm_stream) cudaStreamDestroy(r->comm_stream);
  cudaGetLastError();
  if (prev >= 0 && prev != r->device) cudaSetDevice(prev);
  delete r;
}

void mfa_ring_set_reserved_sms(mfa_ring_t ring, int32_t sms) { if (ring) reinterpret_cast<Ring*>(ring)->reserve_sms = sms < 1 ? 1 : sms; }
uint64_t mfa_ring_launch_count(mfa_ring_t ring) { return ring ? reinterpret_cast<Ring*>(ring)->launches : 0; }

// q, k, v: this rank's [low | high] chunk pair, device-resident, contiguous [B, H, 2C, D] in `precision` (bf16 / fp16).
// out: fp32 [B, H, 2C, D], lse: fp32 [B, H, 2C] (log2 units).  The attention launch is enqueued on `stream` (the exchange on
// the ring's own side stream); the call returns without synchronising unless stream is NULL.
mfa_error_t mfa_ring_attention_forward(mfa_ring_t ring, mfa_buffer_t q, mfa_buffer_t k, mfa_buffer_t v, mfa_buffer_t out,
                                       mfa_buffer_t lse, uint32_t batch_size, uint32_t chunk_rows, uint32_t num_heads,
                                       uint16_t head_dim, float softmax_scale, mfa_precision_t precision, void* stream) {
  if (!ring || !q || !k || !v || !out || !lse) return MFA_ERROR_INVALID_ARGS;
  if (precision != MFA_PRECISION_BF16 && precision != MFA_PRECISION_FP16) return MFA_ERROR_INVALID_ARGS;
  Ring* r = reinterpret_cast<Ring*>(ring);
  std::lock_guard<std::mutex> lock(r->mu);
  const size_t n = (size_t)batch_size * num_heads * 2 * chunk_rows * head_dim, esz = 2;
  if (n == 0) return MFA_SUCCESS;
  void* kd = mfa_buffer_contents(k);
  void* vd = mfa_buffer_contents(v);
  if (!kd || !vd) return MFA_ERROR_INVALID_ARGS;
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  if (prev_dev != r->device) cudaSetDevice(r->device);
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev != r->device ? prev_dev : -1};
  cudaStream_t cs = reinterpret_cast<cudaStream_t>(stream);
  const bool blocking = stream == nullptr;
  cudaStream_t own = nullptr;
  if (blocking) { if (cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking) != cudaSuccess) return MFA_ERROR_EXECUTION_FAILED; cs = own; }
  mfa_error_t rc = MFA_SUCCESS;
  const int W = r->world;
  if (W > 1) {
    const size_t need = (size_t)(W - 1) * n * esz;
    if (r->visit_cap < need) {
      if (r->have_done) cudaEventSynchronize(r->ev_done);
      if (r->k_visit) cudaFree(r->k_visit);
      if (r->v_visit) cudaFree(r->v_visit);
      r->k_visit = r->v_visit = nullptr; r->visit_cap = 0;
      if (cudaMalloc(&r->k_visit, need) != cudaSuccess || cudaMalloc(&r->v_visit, need) != cudaSuccess) {
        cudaGetLastError();
        if (own) cudaStreamDestroy(own);
        return MFA_ERROR_MEMORY_ALLOCATION;
      }
      r->visit_cap = need;
    }
    ++r->epoch;
    // ---- the exchange: step s sends the rank's pair to rank + s and receives slot s from rank - s
    cudaEventRecord(r->ev_ready, cs);                          // the caller's K / V were produced on the compute stream
    cudaStreamWaitEvent(r->comm_stream, r->ev_ready, 0);
    if (r->have_done) cudaStreamWaitEvent(r->comm_stream, r->ev_done, 0);     // the previous forward still reads the slots
    NcclApi& nc = nccl();
    for (int s = 1; s < W; ++s) {
      const int dst = (r->rank + s) % W, src = (r->rank - s + W) % W;
      char* nk = reinterpret_cast<char*>(r->k_visit) + (size_t)(s - 1) * n * esz;
      char* nv = reinterpret_cast<char*>(r->v_visit) + (size_t)(s - 1) * n * esz;
      ncclResult_t e = nc.GroupStart();
      if (e == 0) e = nc.Send(kd, n * esz, 0, dst, r->comm, r->comm_stream);
      if (e == 0) e = nc.Send(vd, n * esz, 0, dst, r->comm, r->comm_stream);
      if (e == 0) e = nc.Recv(nk, n * esz, 0, src, r->comm, r->comm_stream);
      if (e == 0) e = nc.Recv(nv, n * esz, 0, src, r->comm, r->comm_stream);
      const ncclResult_t e2 = nc.GroupEnd();
      if (e != 0 || e2 != 0) {
        RDBG("nccl send/recv: %s", nc.GetErrorString ? nc.GetErrorString(e ? e : e2) : "error");
        rc = MFA_ERROR_EXECUTION_FAILED;
        break;
      }
      if (StreamWriteValue32Fn wv = stream_write_value32()) {
        if (wv(r->comm_stream, reinterpret_cast<CUdeviceptr>(r->flags + s), r->epoch, 0) != CUDA_SUCCESS) rc = MFA_ERROR_EXECUTION_FAILED;
      } else {
        ring_raise_flag_kernel<<<1, 1, 0, r->comm_stream>>>(r->flags + s, r->epoch);
        if (cudaGetLastError() != cudaSuccess) rc = MFA_ERROR_EXECUTION_FAILED;
      }
      if (rc != MFA_SUCCESS) break;
    }
  }
  // ---- the computation: one persistent launch over the rank's own pair and every slot, in arrival order
  if (rc == MFA_SUCCESS) {
    rc = mfa_attention_forward_ring_slots(r->ctx, q, k, v, out, lse, r->k_visit, r->v_visit, r->flags, r->epoch, r->rank, W,
                                          batch_size, chunk_rows, num_heads, head_dim, softmax_scale, precision,
                                          W > 1 ? r->reserve_sms : 0, cs);
    ++r->launches;
  }
  if (rc == MFA_SUCCESS && W > 1) { cudaEventRecord(r->ev_done, cs); r->have_done = true; }
  if (rc != MFA_SUCCESS && W > 1) {
    // never leave a launched kernel polling for slots that will not come: raise every flag
    for (int s = 1; s < W; ++s) ring_raise_flag_kernel<<<1, 1, 0, r->comm_stream>>>(r->flags + s, r->epoch);
    cudaGetLastError();
  }
  if (blocking) {
    if (cudaStreamSynchronize(cs) != cudaSuccess) { cudaGetLastError(); rc = rc == MFA_SUCCESS ? MFA_ERROR_EXECUTION_FAILED : rc; }
    cudaStreamDestroy(own);
  }
  return rc;
}

}  // extern "C"
