/*
 * mfa_ffi.h -- C ABI of libMFAFFI.so, the B200 (sm_100a) flash-attention engine.
 *
 * Binary-compatible with the reference's Sources/MFAFFI/include/mfa_ffi.h: every enum value, struct
 * layout, symbol name and argument list below equals the reference's, so its Rust (bindgen),
 * ctypes, Objective-C and PyTorch C++ adapters link unchanged.  Each declaration cites the reference
 * line it replaces ("ref:" = Sources/MFAFFI/include/mfa_ffi.h, "bridge:" = Sources/MFABridge/MFABridge.swift).
 * Symbols the reference exports without declaring them, and the additive B200 symbols
 * (sliding window, CUDA streams, multi-GPU), live in mfa_ffi_ext.h.
 *
 * B200 semantics that differ from a unified-memory Mac are called out per function.  Tensors are
 * row-major BHSD ([batch, heads, seq, head_dim]) -- what the reference kernel indexes
 * (AttentionKernel+Source.swift:104-121) and every shipped adapter passes, whatever ref:252-255 says.
 */
#ifndef MFA_FFI_H
#define MFA_FFI_H

#ifdef __cplusplus
extern "C" {
#endif

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

/* ref:17-26 -- 0 on success, small positive codes on failure; entry points never abort the process. */
enum {
    MFA_SUCCESS = 0,
    MFA_ERROR_INVALID_ARGS = 1,
    MFA_ERROR_MEMORY_ALLOCATION = 2,
    MFA_ERROR_DEVICE_NOT_SUPPORTED = 3,   /* no sm_100 device visible */
    MFA_ERROR_KERNEL_COMPILATION = 4,     /* kept for ABI; kernels are precompiled, never returned */
    MFA_ERROR_EXECUTION_FAILED = 5
};
typedef int mfa_error_t;

/* ref:33-41 */
enum {
    MFA_PRECISION_FP16 = 0,
    MFA_PRECISION_BF16 = 1,
    MFA_PRECISION_FP32 = 2,
    MFA_PRECISION_INT8 = 3,
    MFA_PRECISION_INT4 = 4
};
typedef int mfa_precision_t;

/* ref:46-52 */
enum { MFA_MASK_TYPE_NONE = 0, MFA_MASK_TYPE_BOOL = 1, MFA_MASK_TYPE_ADDITIVE = 2 };
typedef int mfa_mask_type_t;

/* ref:57-64 */
enum { MFA_MASK_SCALAR_BYTE = 0, MFA_MASK_SCALAR_FP16 = 1, MFA_MASK_SCALAR_BF16 = 2, MFA_MASK_SCALAR_FP32 = 3 };
typedef int mfa_mask_scalar_t;

typedef void* mfa_context_t;   /* ref:69 -- CUDA device + stream + scratch, process-wide singleton */
typedef void* mfa_buffer_t;    /* ref:74 -- host/device memory view */

/* ref:76-80 */
typedef enum {
    MFA_QUANT_KERNEL_FORWARD = 0,
    MFA_QUANT_KERNEL_BACKWARD_QUERY = 1,
    MFA_QUANT_KERNEL_BACKWARD_KEY_VALUE = 2
} mfa_quantized_kernel_t;

/* ref:82-121 -- Metal argument-table slots.  Meaningless on CUDA; the reference's own FFI fills every
 * field with -1 (QuantizedLayoutManifest+FFI.swift:146-155) and so does this library. */
typedef struct {
    int32_t qData, kData, vData, output, gradOutput, logsumexp, gradQuery, dValues, gradKey, gradValue;
    int32_t qScale, qZeroPoint, kScale, kZeroPoint, vScale, vZeroPoint;
    int32_t dims, steClipRange;
    int32_t qBlockScales, qBlockZeroPoints, kBlockScales, kBlockZeroPoints, vBlockScales, vBlockZeroPoints;
    int32_t qPrecomputedSums, kPrecomputedSums, vPrecomputedSums;
    int32_t qStrides, kStrides, vStrides, oStrides;
    int32_t maskBuffer, numHeads, numKeyValueHeads, headDimension, sequenceLength;
    int32_t scratch0, scratch1;
} mfa_quantized_layout_t;

void mfa_get_quantized_layout(mfa_quantized_kernel_t kernel, mfa_quantized_layout_t* out_layout);  /* ref:123-126 */

/* ref:128-133 */
typedef struct {
    bool supports_multi_head_backward;
    bool supports_blockwise_backward;
    uint32_t max_heads;
    uint32_t max_block_size;
} mfa_quantized_capabilities_t;

void mfa_get_quantized_capabilities(void* out_capabilities);   /* ref:135 */

/* ---- context (ref:147,154; bridge:782-805).  create returns +1 reference on a process-wide singleton,
 * destroy drops one; any number of create/destroy pairs is safe. */
mfa_error_t mfa_create_context(mfa_context_t* context);
void mfa_destroy_context(mfa_context_t context);

/* ---- buffers (ref:168-239; bridge:850-1063).
 * create_buffer      : library-owned allocation whose mfa_buffer_contents() pointer is CPU-dereferenceable
 *                      (pinned host memory mirrored in HBM; Metal "shared" storage equivalent).
 * buffer_from_ptr    : wraps caller memory, never owns or frees it.  A device or managed pointer is used in
 *                      place (true zero-copy).  A host pointer gets an HBM mirror: inputs are copied in when a
 *                      compute call starts, outputs copied back before it returns, so results are visible in
 *                      the caller's array on return exactly as on unified memory (bridge:1412-1413).
 * from_mtl_buffer    : `metal_buffer` is a CUDA device pointer (e.g. torch.Tensor.data_ptr()).
 * *_with_strides     : shape/strides are element counts in BHSD order, last dimension contiguous.
 * destroy_buffer     : frees the handle (and the allocation only if the library made it). */
mfa_error_t mfa_create_buffer(mfa_context_t context, size_t size_bytes, mfa_buffer_t* buffer);
mfa_error_t mfa_buffer_from_ptr(mfa_context_t context, void* data_ptr, size_t size_bytes, mfa_buffer_t* buffer);
mfa_error_t mfa_buffer_from_ptr_with_strides(mfa_context_t context, void* data_ptr, size_t size_bytes,
                                             const int64_t* shape, const int64_t* strides, uint32_t ndim,
                                             mfa_buffer_t* buffer);
mfa_error_t mfa_buffer_from_mtl_buffer(mfa_context_t context, void* metal_buffer, size_t size_bytes,
                                       mfa_buffer_t* buffer);
mfa_error_t mfa_buffer_from_mtl_buffer_with_strides(mfa_context_t context, void* metal_buffer, size_t size_bytes,
                                                    const int64_t* shape, const int64_t* strides, uint32_t ndim,
                                                    mfa_buffer_t* buffer);
void* mfa_buffer_contents(mfa_buffer_t buffer);
void mfa_destroy_buffer(mfa_buffer_t buffer);

/* ---- forward (ref:273-300; bridge:1074-1433).  O = softmax(scale * Q K^T [+ mask]) V, blocking.
 * O is written as fp32 whatever output_precision says (the reference ignores it, bridge:1090) unless the
 * `out` handle is too small for fp32 and exactly fits output_precision, in which case that type is written.
 * causal: key j visible to query i iff j <= i (top-left aligned).  mask: BOOL (non-zero byte = attend) or
 * ADDITIVE (fp16/bf16/fp32), up to 4-D, right-aligned broadcast to [B,H,Sq,Skv]; added after scaling
 * (PyTorch semantics -- SURVEY quirk Q7).  Returns 1 on NULL handles / bad sizes, 5 on CUDA failure. */
mfa_error_t mfa_attention_forward(
    mfa_context_t context, mfa_buffer_t q, mfa_buffer_t k, mfa_buffer_t v, mfa_buffer_t out,
    uint32_t batch_size, uint32_t seq_len_q, uint32_t seq_len_kv, uint32_t num_heads, uint16_t head_dim,
    float softmax_scale, bool causal,
    mfa_precision_t input_precision, mfa_precision_t intermediate_precision, mfa_precision_t output_precision,
    bool transpose_q, bool transpose_k, bool transpose_v, bool transpose_o,
    const void* mask_ptr, size_t mask_size_bytes, const int64_t* mask_shape, const int64_t* mask_strides,
    uint32_t mask_ndim, mfa_mask_type_t mask_type, mfa_mask_scalar_t mask_scalar_type);

/* ref:312-334; bridge:2377-2543.  Asynchronous variant: `command_buffer` is a cudaStream_t, buffers are raw
 * device pointers with byte offsets, strides are BHSD element strides (NULL = contiguous).  Enqueues and
 * returns without synchronising -- the caller owns completion, as with an MTLCommandBuffer. */
mfa_error_t mfa_attention_encode_mtl(
    mfa_context_t context, void* command_buffer,
    void* q_buffer, int64_t q_offset, const int64_t* q_strides,
    void* k_buffer, int64_t k_offset, const int64_t* k_strides,
    void* v_buffer, int64_t v_offset, const int64_t* v_strides,
    void* out_buffer, int64_t out_offset,
    void* mask_buffer, int64_t mask_offset, const int64_t* mask_shape, const int64_t* mask_strides,
    uint32_t mask_ndim, mfa_mask_type_t mask_type, mfa_mask_scalar_t mask_scalar_type,
    uint32_t batch_size, uint32_t seq_len_q, uint32_t seq_len_kv, uint32_t num_heads, uint16_t head_dim,
    float softmax_scale, bool causal, const char* input_precision, const char* intermediate_precision);

/* ref:364-391; bridge:2671-2899.  Quantised forward.  Precision arguments say what the buffers hold:
 *   q/k/v_precision in {FP16,BF16,FP32}: the operand is quantised on the device (symmetric, per-tensor
 *     absmax/127 or /7 -- GEMMRuntimeQuantization.swift:80-181) to the integer width requested by the other
 *     operands, or left as is if none is integer;
 *   q/k/v_precision in {INT8,INT4}: the buffer already holds codes (int4 packed low nibble first) and
 *     x_scale / x_zero_point dequantise them as (code - zero_point) * scale.
 * The reference ignores all of these arguments (MFABridge+Quantized.swift:26-35); this library honours them. */
mfa_error_t mfa_attention_forward_quantized(
    mfa_context_t context, mfa_buffer_t q, mfa_buffer_t k, mfa_buffer_t v, mfa_buffer_t out,
    uint32_t batch_size, uint32_t seq_len_q, uint32_t seq_len_kv, uint32_t num_heads, uint16_t head_dim,
    float softmax_scale, bool causal,
    float q_scale, int32_t q_zero_point, float k_scale, int32_t k_zero_point, float v_scale, int32_t v_zero_point,
    mfa_precision_t q_precision, mfa_precision_t k_precision, mfa_precision_t v_precision,
    mfa_precision_t output_precision,
    bool transpose_q, bool transpose_k, bool transpose_v, bool transpose_o);

/* ref:393-405 -- DeepSeek indexer GEMM (MPS on the reference).  Out of the hot path: returns
 * MFA_ERROR_DEVICE_NOT_SUPPORTED. */
mfa_error_t mfa_sparse_indexer_scores(
    mfa_context_t context, mfa_buffer_t q, mfa_buffer_t k,
    uint32_t batch_size, uint32_t num_heads, uint32_t seq_len_q, uint32_t seq_len_k, uint16_t head_dim,
    float scale, mfa_buffer_t scores_in, mfa_buffer_t* scores_out);

/* ref:407-438; bridge:3171-3282.  dQ, dK, dV (fp32) from dO, Q, K, V, O (fp32) and L.
 * softmax_lse is the forward's L = log2(e) * logsumexp(scale * S) (fp32 [B,H,Sq]); d_buffer receives
 * D = scale * rowsum(dO * O) (fp32 [B,H,Sq], zeroed and filled by the callee). */
mfa_error_t mfa_attention_backward(
    mfa_context_t context,
    mfa_buffer_t dout, mfa_buffer_t q, mfa_buffer_t k, mfa_buffer_t v, mfa_buffer_t out, mfa_buffer_t softmax_lse,
    mfa_buffer_t dq, mfa_buffer_t dk, mfa_buffer_t dv, mfa_buffer_t d_buffer,
    uint32_t batch_size, uint32_t seq_len_q, uint32_t seq_len_kv, uint32_t num_heads, uint16_t head_dim,
    float softmax_scale, bool causal,
    mfa_precision_t input_precision, mfa_precision_t intermediate_precision,
    bool transpose_q, bool transpose_k, bool transpose_v, bool transpose_o);

/* ---- utilities (ref:450-478; bridge:1528-1617) */
const char* mfa_error_string(mfa_error_t error);        /* strdup'd: caller frees with free() */
bool mfa_is_device_supported(void);                     /* true iff an sm_100 GPU is visible */
void mfa_get_version(int* major, int* minor, int* patch);   /* 1.0.0 */
double mfa_get_gpu_latency(mfa_context_t context);      /* seconds, CUDA-event time of the last blocking op */

/* ---- quantised backward on pre-quantised operands (ref:480-624; bridge:1699-2163).  q/k/v hold int8/int4
 * codes with per-tensor scale/zero-point (and, for *_ex, optional per-block fp32 scales / int32 zero points
 * over blocks of <x>_block_size tokens); precisions use THIS header's enum. */
int32_t mfa_attention_backward_query_quantized(
    mfa_context_t context, mfa_buffer_t q, mfa_buffer_t k, mfa_buffer_t v, mfa_buffer_t output,
    mfa_buffer_t grad_output, mfa_buffer_t logsumexp, mfa_buffer_t grad_query, mfa_buffer_t d_values,
    uint32_t batch_size, uint32_t seq_len_q, uint32_t seq_len_kv, uint32_t num_heads, uint16_t head_dim,
    float q_scale, int32_t q_zero_point, float k_scale, int32_t k_zero_point, float v_scale, int32_t v_zero_point,
    int32_t q_precision, int32_t k_precision, int32_t v_precision, bool causal,
    bool transpose_q, bool transpose_k, bool transpose_v, bool transpose_o);

int32_t mfa_attention_backward_kv_quantized(
    mfa_context_t context, mfa_buffer_t q, mfa_buffer_t k, mfa_buffer_t v,
    mfa_buffer_t grad_output, mfa_buffer_t logsumexp, mfa_buffer_t d_values,
    mfa_buffer_t grad_key, mfa_buffer_t grad_value,
    uint32_t batch_size, uint32_t seq_len_q, uint32_t seq_len_kv, uint32_t num_heads, uint16_t head_dim,
    float q_scale, int32_t q_zero_point, float k_scale, int32_t k_zero_point, float v_scale, int32_t v_zero_point,
    int32_t q_precision, int32_t k_precision, int32_t v_precision, bool causal,
    bool transpose_q, bool transpose_k, bool transpose_v, bool transpose_o);

int32_t mfa_attention_backward_query_quantized_ex(
    mfa_context_t context, mfa_buffer_t q, mfa_buffer_t k, mfa_buffer_t v, mfa_buffer_t output,
    mfa_buffer_t grad_output, mfa_buffer_t logsumexp, mfa_buffer_t grad_query, mfa_buffer_t d_values,
    uint32_t batch_size, uint32_t seq_len_q, uint32_t seq_len_kv, uint32_t num_heads, uint32_t num_kv_heads,
    uint16_t head_dim,
    float q_scale, int32_t q_zero_point, float k_scale, int32_t k_zero_point, float v_scale, int32_t v_zero_point,
    int32_t q_precision, int32_t k_precision, int32_t v_precision, bool causal,
    bool transpose_q, bool transpose_k, bool transpose_v, bool transpose_o,
    mfa_buffer_t q_block_scales, mfa_buffer_t q_block_zero_points,
    mfa_buffer_t k_block_scales, mfa_buffer_t k_block_zero_points,
    mfa_buffer_t v_block_scales, mfa_buffer_t v_block_zero_points,
    uint32_t q_block_size, uint32_t k_block_size, uint32_t v_block_size, uint32_t options);

int32_t mfa_attention_backward_kv_quantized_ex(
    mfa_context_t context, mfa_buffer_t q, mfa_buffer_t k, mfa_buffer_t v,
    mfa_buffer_t grad_output, mfa_buffer_t logsumexp, mfa_buffer_t d_values,
    mfa_buffer_t grad_key, mfa_buffer_t grad_value,
    uint32_t batch_size, uint32_t seq_len_q, uint32_t seq_len_kv, uint32_t num_heads, uint32_t num_kv_heads,
    uint16_t head_dim,
    float q_scale, int32_t q_zero_point, float k_scale, int32_t k_zero_point, float v_scale, int32_t v_zero_point,
    int32_t q_precision, int32_t k_precision, int32_t v_precision, bool causal,
    bool transpose_q, bool transpose_k, bool transpose_v, bool transpose_o,
    mfa_buffer_t q_block_scales, mfa_buffer_t q_block_zero_points,
    mfa_buffer_t k_block_scales, mfa_buffer_t k_block_zero_points,
    mfa_buffer_t v_block_scales, mfa_buffer_t v_block_zero_points,
    uint32_t q_block_size, uint32_t k_block_size, uint32_t v_block_size, uint32_t options);

/* ---- MLA latent-KV decompression (ref:644-720).  A separate GEMM feature outside the attention hot path
 * (SURVEY section 2 row 9): symbols exist so the header links; create/destroy work, the compute entry
 * points return MFA_ERROR_DEVICE_NOT_SUPPORTED. */
typedef void* mfa_mla_context_t;
mfa_error_t mfa_mla_create_context(mfa_mla_context_t* context);
void mfa_mla_destroy_context(mfa_mla_context_t context);
mfa_error_t mfa_mla_init_weights(mfa_mla_context_t context, uint32_t num_heads, uint32_t head_dim,
                                 uint32_t kv_latent_dim);
mfa_error_t mfa_mla_load_weights(mfa_mla_context_t context, mfa_buffer_t wk, mfa_buffer_t wv);
mfa_error_t mfa_mla_forward(mfa_mla_context_t context, mfa_context_t mfa_context, mfa_buffer_t kv_latent,
                            mfa_buffer_t* decompressed_k, mfa_buffer_t* decompressed_v,
                            uint32_t batch_size, uint32_t num_heads, uint32_t sequence_length,
                            uint32_t head_dim, uint32_t kv_latent_dim);

#ifdef __cplusplus
}
#endif

#endif /* MFA_FFI_H */
