// common.h -- shared host/device definitions for libMFAFFI.so (B200 / sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mfa {

// Element types, numbered like mfa_precision_t (include/mfa_ffi.h).
enum DType : int { kF16 = 0, kBF16 = 1, kF32 = 2, kI8 = 3, kI4 = 4 };

inline __host__ __device__ int dtype_bytes(int dt) { return dt == kF32 ? 4 : (dt == kI8 || dt == kI4) ? 1 : 2; }

enum MaskKind : int { kMaskNone = 0, kMaskBool = 1, kMaskAdditive = 2 };
enum MaskScalar : int { kMaskU8 = 0, kMaskF16 = 1, kMaskBF16 = 2, kMaskF32 = 3 };

// A [B, H, S, D] operand view: element (b,h,s,d) lives at ptr + b*sb + h*sh + s*ss + d*sd (element units).
// Contiguous BHSD has sd = 1, ss = D, sh = S*D, sb = H*S*D; transpose_x (reference: [D,S] per head) has
// ss = 1, sd = S.
struct TensorView {
  const void* ptr;
  int64_t sb, sh, ss, sd;
};

// Per-tensor / per-block symmetric quantisation metadata for an int8/int4 operand.
// block_rows tokens share one scale (0 = one scale for the whole tensor); scales laid out [B*H*ceil(S/block_rows)]
// when block_rows > 0.  value = (code - zero_point) * scale.
struct QuantView {
  const float* scales;   // device pointer, or nullptr -> use `scale`
  float scale;
  int zero_point;
  int block_rows;
};

struct AttnParams {
  TensorView q, k, v;       // inputs, in_dtype
  TensorView o;             // output (forward) or saved O (backward), o_dtype
  float* lse;               // [B,H,Sq] fp32, log2 units (L = log2e * logsumexp); may be null in forward
  int64_t lse_sh;           // forward, tensor-core path: elements between L rows of consecutive (b,h) (0 = Sq)
  int accumulate;           // forward, tensor-core path: merge the result into the (O, L) already in o / lse
  // backward only
  TensorView d_o;           // dO, do_dtype
  float* dq;                // fp32, contiguous BHSD
  float* dk;
  float* dv;
  float* dterm;             // [B,H,Sq] fp32: scale * rowsum(dO*O)
  int B, H, Hkv, Sq, Skv, D;
  float scale;
  int causal;               // key j visible to query i iff j <= i
  int window;               // <0: none; else hidden iff i > j + window
  // external mask, broadcast via zero strides
  const void* mask;
  int mask_kind, mask_scalar;
  int64_t mask_sb, mask_sh, mask_sq, mask_sk;
  // tensor-core forward: device scratch for the per-query-block lists of KV tiles the mask leaves visible (tile skipping);
  // null = every tile in the causal / window range is visited.  Size: fwd_tc_mask_scratch_bytes(p).
  int* mask_tile_scratch;
  int in_dtype, o_dtype, do_dtype;
  QuantView qq, qk, qv;     // used when in_dtype is kI8 / kI4
};

// Visible key range [lo, hi) for query rows [r0, r1) under causal/window rules (used for tile skipping).
__host__ __device__ inline void visible_key_range(int causal, int window, int Skv, int r0, int r1, int& lo, int& hi) {
  lo = 0; hi = Skv;
  if (causal) { int h = r1; if (h < hi) hi = h; }            // cols <= r1-1
  if (window >= 0) { int l = r0 - window; if (l > lo) lo = l; } // cols >= r0 - window
  if (hi < lo) hi = lo;
}

// Visible query range [lo, hi) for key columns [c0, c1).
__host__ __device__ inline void visible_query_range(int causal, int window, int Sq, int c0, int c1, int& lo, int& hi) {
  lo = 0; hi = Sq;
  if (causal) { if (c0 > lo) lo = c0; }                        // rows >= c0
  if (window >= 0) { long h = (long)c1 - 1 + window + 1; if (h < hi) hi = (int)h; } // rows <= c1-1+window
  if (hi < lo) hi = lo;
}

// ---- kernel launchers (each returns cudaGetLastError()) ----
cudaError_t launch_fwd_simt(const AttnParams& p, cudaStream_t st);
cudaError_t launch_bwd_simt(const AttnParams& p, cudaStream_t st);          // dterm + dQ + dK/dV
cudaError_t launch_dterm(const AttnParams& p, cudaStream_t st);
cudaError_t launch_bwd_dkv_only(const AttnParams& p, cudaStream_t st);      // dK/dV from a caller-supplied dterm

// tcgen05 forward (bf16/fp16, D in {64,128}); returns cudaErrorNotSupported when the problem is not eligible.
bool fwd_tc_eligible(const AttnParams& p);
size_t fwd_tc_mask_scratch_bytes(const AttnParams& p);     // 0 when there is no external mask
size_t bwd_tc_mask_scratch_bytes(const AttnParams& p);     // same for the backward's two list sets
void mask_tile_dims(const AttnParams& p, int& nqb, int& nkt, int& MB, int& MH);
cudaError_t launch_mask_flags(const AttnParams& p, uint8_t* flags, cudaStream_t st);
cudaError_t launch_fwd_tc(const AttnParams& p, cudaStream_t st);

// tcgen05 backward (bf16/fp16, D in {64,128}); needs p.dterm filled by launch_dterm first.
bool bwd_tc_eligible(const AttnParams& p);
cudaError_t launch_bwd_tc(const AttnParams& p, cudaStream_t st);

// fp32 operands on the tensor pipe as fp16 (hi, lo) pairs (attn_fwd_split.cu): head_dim 128
bool fwd_split_eligible(const AttnParams& p);
size_t fwd_split_scratch_bytes(const AttnParams& p);
cudaError_t launch_fwd_split(const AttnParams& p, void* scratch, cudaStream_t st);

// int8 tensor-core forward (attn_fwd_tcq.cu): int8 / int4 codes, symmetric, D = 128, per-tensor or 64-multiple block scales.
bool fwd_tcq_eligible(const AttnParams& p);
size_t fwd_tcq_scratch_bytes(const AttnParams& p);
cudaError_t launch_fwd_tcq(const AttnParams& p, void* scratch, cudaStream_t st);
cudaError_t launch_codes_to_bf16(const void* codes, int bits, void* dst, uint64_t n, cudaStream_t st);
int fwd_tcq_pv_mode();                      // P V of the quantised tensor-core forward: 0 = e4m3 (default), 1 = bf16
void fwd_tcq_set_pv_mode(int bf16);

// quantiser & friends
cudaError_t launch_quantize(const void* src, int src_dtype, void* codes, float* scales, uint64_t rows, uint64_t cols,
                            uint32_t block_rows, uint32_t block_cols, int bits, float scale_floor, cudaStream_t st);
cudaError_t launch_dequantize(const void* codes, const float* scales, float* out, uint64_t rows, uint64_t cols,
                              uint32_t block_rows, uint32_t block_cols, int bits, cudaStream_t st);
// tensor-core backward of quantised operands: dequantised codes / the upstream gradient as bf16
cudaError_t launch_dequantize_bf16(const void* codes, int bits, const QuantView& q, void* out, uint64_t heads,
                                   uint64_t rows_per_head, uint32_t D, cudaStream_t st);
cudaError_t launch_to_bf16(const void* src, int src_dtype, void* out, uint64_t n, cudaStream_t st);
cudaError_t launch_merge_partials(float* o_acc, float* l_acc, const float* o_part, const float* l_part,
                                  uint64_t rows, uint32_t D, cudaStream_t st);
cudaError_t launch_hadamard(float* data, uint32_t block_size, uint32_t num_blocks, cudaStream_t st);
cudaError_t launch_rope(const void* src, void* dst, const float* cos_t, const float* sin_t, int64_t sB, int64_t sH,
                        int64_t sS, int64_t table_batch_stride, bool negate_sin, uint32_t B, uint32_t H, uint32_t S,
                        uint32_t D, int dtype, cudaStream_t st);
cudaError_t launch_convert_from_f32(const float* src, void* dst, int dst_dtype, uint64_t n, cudaStream_t st);

// Count of kernel launches performed by this library (bench.py's gpu_launches claim).
extern unsigned long long g_launch_count;
extern const char* g_last_kernel;

}  // namespace mfa
