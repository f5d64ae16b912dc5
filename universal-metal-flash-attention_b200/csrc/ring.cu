// ring.cu -- native ring attention (context parallelism over the GPUs of one node): the mfa_ring_* symbols of
// include/mfa_ffi_ext.h.  No reference counterpart ("Reference support: none"); this is BASELINE.json's
// north_star item (4): "for long causal contexts, ring attention that passes K/V blocks over NVLink with NCCL send/recv
// overlapped with compute".
//
// One process per GPU.  The sequence is cut into 2*world chunks, rank r owns chunks r and 2*world-1-r (zig-zag, so causal work is
// balanced) stored next to each other: q / k / v are [B, H, 2C, D] with the low chunk first.  Q stays put, the K/V pair travels
// round the ring; with that layout the visible chunk pairs of a step are ONE rectangular problem (step_plan), i.e. one kernel
// launch per step through the library's own C entry points:
//     step 0           mfa_attention_forward_ex            causal 2C x 2C on local indices
//     source < rank    mfa_attention_forward_accumulate    2C x C  (both local query chunks see the visitor's low chunk)
//     source > rank    mfa_attention_forward_accumulate     C x 2C (the high query chunk sees both chunks of the visitor)
// and the partial (O, L) of a step is merged into the running result inside the attention epilogue.
//
// Transport: ncclSend / ncclRecv of the K and V halves, grouped, on a highest-priority side stream, double-buffered so hop s+1
// overlaps the attention of hop s.  NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the host process already
// loaded -- e.g. PyTorch's -- or the system one), so libMFAFFI.so keeps loading on machines without NCCL.
// The attention grid is persistent (one CTA per SM), so an SM-resident NCCL kernel would only be scheduled when the grid ends:
// while a hop is in flight the step's attention launch leaves `reserve_sms` SMs to the transport (mfa_ring_set_reserved_sms).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/mfa_ffi_ext.h"

namespace mfa {
void fwd_tc_set_sm_limit(int sms);      // attn_fwd_tc.cu: cap the persistent grid of the next launches (0 = all SMs)
}

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

struct Ring {
  mfa_context_t ctx = nullptr;
  int rank = 0, world = 1, device = 0;
  ncclComm_t comm = nullptr;
  bool own_comm = false;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_arrived[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  void* kv[2] = {nullptr, nullptr};        // visiting K/V pairs: [K | V], each B*H*2C*D elements
  size_t kv_cap = 0;
  int reserve_sms = 8;
  unsigned long long launches = 0;
  std::mutex mu;
};

bool debug_on() { const char* d = getenv("MFA_DEBUG"); return d && d[0] && d[0] != '0'; }
#define RDBG(...) do { if (debug_on()) { fprintf(stderr, "[mfa ring] " __VA_ARGS__); fputc('\n', stderr); } } while (0)

// the rectangular problem of ring step `step` on rank `rank` (umfa/ring.py step_plan; units of chunks)
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

static mfa_error_t ring_finish_create(Ring* r, mfa_ring_t* out) {
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);                  // hi = numerically lowest = highest priority
  if (cudaStreamCreateWithPriority(&r->comm_stream, cudaStreamNonBlocking, hi) != cudaSuccess) { cudaGetLastError(); delete r; return MFA_ERROR_EXECUTION_FAILED; }
  bool ok = cudaEventCreateWithFlags(&r->ev_ready, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < 2; ++i) {
    ok = ok && cudaEventCreateWithFlags(&r->ev_arrived[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&r->ev_consumed[i], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) { cudaGetLastError(); delete r; return MFA_ERROR_EXECUTION_FAILED; }
  if (const char* e = getenv("MFA_RING_RESERVE_SMS")) r->reserve_sms = atoi(e);
  *out = reinterpret_cast<mfa_ring_t>(r);
  return MFA_SUCCESS;
}

mfa_error_t mfa_ring_create(mfa_context_t context, const void* unique_id, size_t id_bytes, int32_t rank, int32_t world_size,
                            mfa_ring_t* ring) {
  if (!context || !ring || rank < 0 || world_size < 1 || rank >= world_size) return MFA_ERROR_INVALID_ARGS;
  *ring = nullptr;
  Ring* r = new (std::nothrow) Ring();
  if (!r) return MFA_ERROR_MEMORY_ALLOCATION;
  r->ctx = context; r->rank = rank; r->world = world_size;
  cudaGetDevice(&r->device);
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
  if (!context || !ring || rank < 0 || world_size < 1 || rank >= world_size) return MFA_ERROR_INVALID_ARGS;
  *ring = nullptr;
  if (world_size > 1 && (!nccl_comm || !nccl().ok)) return nccl_comm ? MFA_ERROR_DEVICE_NOT_SUPPORTED : MFA_ERROR_INVALID_ARGS;
  Ring* r = new (std::nothrow) Ring();
  if (!r) return MFA_ERROR_MEMORY_ALLOCATION;
  r->ctx = context; r->rank = rank; r->world = world_size;
  r->comm = reinterpret_cast<ncclComm_t>(nccl_comm);
  cudaGetDevice(&r->device);
  return ring_finish_create(r, ring);
}

void mfa_ring_destroy(mfa_ring_t ring) {
  if (!ring) return;
  Ring* r = reinterpret_cast<Ring*>(ring);
  if (r->comm_stream) cudaStreamSynchronize(r->comm_stream);
  if (r->own_comm && r->comm) nccl().CommDestroy(r->comm);
  for (int i = 0; i < 2; ++i) {
    if (r->kv[i]) cudaFree(r->kv[i]);
    if (r->ev_arrived[i]) cudaEventDestroy(r->ev_arrived[i]);
    if (r->ev_consumed[i]) cudaEventDestroy(r->ev_consumed[i]);
  }
  if (r->ev_ready) cudaEventDestroy(r->ev_ready);
  if (r->comm_stream) cudaStreamDestroy(r->comm_stream);
  cudaGetLastError();
  delete r;
}

void mfa_ring_set_reserved_sms(mfa_ring_t ring, int32_t sms) { if (ring) reinterpret_cast<Ring*>(ring)->reserve_sms = sms < 0 ? 0 : sms; }
uint64_t mfa_ring_launch_count(mfa_ring_t ring) { return ring ? reinterpret_cast<Ring*>(ring)->launches : 0; }

// q, k, v: this rank's [low | high] chunk pair, device-resident, contiguous [B, H, 2C, D] in `precision` (bf16 / fp16).
// out: fp32 [B, H, 2C, D], lse: fp32 [B, H, 2C] (log2 units).  Everything is enqueued on `stream` (the compute stream; hops on
// the ring's own side stream) and the call returns without synchronising unless stream is NULL.
mfa_error_t mfa_ring_attention_forward(mfa_ring_t ring, mfa_buffer_t q, mfa_buffer_t k, mfa_buffer_t v, mfa_buffer_t out,
                                       mfa_buffer_t lse, uint32_t batch_size, uint32_t chunk_rows, uint32_t num_heads,
                                       uint16_t head_dim, float softmax_scale, mfa_precision_t precision, void* stream) {
  if (!ring || !q || !k || !v || !out || !lse) return MFA_ERROR_INVALID_ARGS;
  if (precision != MFA_PRECISION_BF16 && precision != MFA_PRECISION_FP16) return MFA_ERROR_INVALID_ARGS;
  Ring* r = reinterpret_cast<Ring*>(ring);
  std::lock_guard<std::mutex> lock(r->mu);
  const int64_t B = batch_size, H = num_heads, C = chunk_rows, T = 2 * C, D = head_dim;
  if (B == 0 || H == 0 || C == 0 || D == 0) return MFA_SUCCESS;
  const size_t n = (size_t)B * H * T * D, esz = 2;
  void* qd = mfa_buffer_contents(q);
  void* kd = mfa_buffer_contents(k);
  void* vd = mfa_buffer_contents(v);
  if (!qd || !kd || !vd) return MFA_ERROR_INVALID_ARGS;
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  if (prev_dev != r->device) cudaSetDevice(r->device);
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev != r->device ? prev_dev : -1};
  cudaStream_t cs = reinterpret_cast<cudaStream_t>(stream);
  const bool blocking = stream == nullptr;
  cudaStream_t own = nullptr;
  if (blocking) { if (cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking) != cudaSuccess) return MFA_ERROR_EXECUTION_FAILED; cs = own; }
  mfa_error_t rc = MFA_SUCCESS;
  if (r->world > 1 && r->kv_cap < 2 * n * esz) {
    for (int i = 0; i < 2; ++i) { if (r->kv[i]) cudaFree(r->kv[i]); r->kv[i] = nullptr; }
    r->kv_cap = 0;
    if (cudaMalloc(&r->kv[0], 2 * n * esz) != cudaSuccess || cudaMalloc(&r->kv[1], 2 * n * esz) != cudaSuccess) {
      cudaGetLastError();
      if (own) cudaStreamDestroy(own);
      return MFA_ERROR_MEMORY_ALLOCATION;
    }
    r->kv_cap = 2 * n * esz;
  }
  const int dst = (r->rank + 1) % r->world, src = (r->rank - 1 + r->world) % r->world;
  char* cur_k = reinterpret_cast<char*>(kd);
  char* cur_v = reinterpret_cast<char*>(vd);
  cudaEventRecord(r->ev_ready, cs);                      // the caller's K / V were produced on the compute stream
  for (int step = 0; step < r->world && rc == MFA_SUCCESS; ++step) {
    const bool hop = step + 1 < r->world;
    if (hop) {
      const int nb = step & 1;
      char* nk = reinterpret_cast<char*>(r->kv[nb]);
      char* nv = nk + n * esz;
      if (step == 0) cudaStreamWaitEvent(r->comm_stream, r->ev_ready, 0);
      if (step >= 2) cudaStreamWaitEvent(r->comm_stream, r->ev_consumed[nb], 0);   // the attention of step-1 read this buffer
      NcclApi& nc = nccl();
      ncclResult_t e = nc.GroupStart();
      if (e == 0) e = nc.Send(cur_k, n * esz, 0, dst, r->comm, r->comm_stream);
      if (e == 0) e = nc.Send(cur_v, n * esz, 0, dst, r->comm, r->comm_stream);
      if (e == 0) e = nc.Recv(nk, n * esz, 0, src, r->comm, r->comm_stream);
      if (e == 0) e = nc.Recv(nv, n * esz, 0, src, r->comm, r->comm_stream);
      const ncclResult_t e2 = nc.GroupEnd();
      if (e != 0 || e2 != 0) { RDBG("nccl send/recv: %s", nc.GetErrorString ? nc.GetErrorString(e ? e : e2) : "error"); rc = MFA_ERROR_EXECUTION_FAILED; break; }
      cudaEventRecord(r->ev_arrived[nb], r->comm_stream);
    }
    const Plan pl = step_plan(r->rank, r->world, step);
    {
      Handle hq(r->ctx, qd, (size_t)pl.q0 * C * D, esz, B, H, (int64_t)pl.qn * C, D, T);
      Handle hk(r->ctx, cur_k, (size_t)pl.k0 * C * D, esz, B, H, (int64_t)pl.kn * C, D, T);
      Handle hv(r->ctx, cur_v, (size_t)pl.k0 * C * D, esz, B, H, (int64_t)pl.kn * C, D, T);
      if (!hq.h || !hk.h || !hv.h) { rc = MFA_ERROR_MEMORY_ALLOCATION; break; }
      mfa::fwd_tc_set_sm_limit(hop && r->reserve_sms > 0 ? -r->reserve_sms : 0);      // leave SMs to the transport while a hop runs
      if (step == 0)
        rc = mfa_attention_forward_ex(r->ctx, hq.h, hk.h, hv.h, out, lse, batch_size, (uint32_t)(pl.qn * C), (uint32_t)(pl.kn * C),
                                      num_heads, head_dim, softmax_scale, pl.causal, -1, precision, MFA_PRECISION_FP32,
                                      nullptr, 0, nullptr, nullptr, 0, MFA_MASK_TYPE_NONE, MFA_MASK_SCALAR_BYTE, cs);
      else
        rc = mfa_attention_forward_accumulate(r->ctx, hq.h, hk.h, hv.h, out, lse, batch_size, (uint32_t)(pl.qn * C),
                                              (uint32_t)(pl.kn * C), num_heads, head_dim, softmax_scale, pl.causal, -1, precision,
                                              (uint32_t)(pl.q0 * C), (uint32_t)T, cs);
      mfa::fwd_tc_set_sm_limit(0);
      ++r->launches;
    }
    if (rc != MFA_SUCCESS) break;
    if (step >= 1) cudaEventRecord(r->ev_consumed[(step - 1) & 1], cs);        // this step's attention read kv[(step-1) & 1]
    if (hop) {
      const int nb = step & 1;
      cudaStreamWaitEvent(cs, r->ev_arrived[nb], 0);
      cur_k = reinterpret_cast<char*>(r->kv[nb]);
      cur_v = cur_k + n * esz;
    }
  }
  if (blocking) {
    if (cudaStreamSynchronize(cs) != cudaSuccess) { cudaGetLastError(); rc = rc == MFA_SUCCESS ? MFA_ERROR_EXECUTION_FAILED : rc; }
    cudaStreamDestroy(own);
  }
  return rc;
}

}  // extern "C"
