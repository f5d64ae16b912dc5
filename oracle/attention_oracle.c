/*
 * attention_oracle.c -- CPU restatement of the reference's attention / quantiser math.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker (or as the timed CPU baseline), never as the thing shipped.
 *
 * What it restates (paths relative to the reference checkout; MFA/ = metal-flash-attention/):
 *   - forward / backward attention math of the reference's own CPU test oracles
 *       MFA/Tests/FlashAttentionTests/Utilities/Network.swift:136-409          (S, P via LSE, L, dP, dS, D, O, dV, dK, dQ)
 *       MFA/Tests/FlashAttentionTests/QuantizedAttentionTest.swift:822-936     (cpuReferenceAttention / cpuReferenceBackward)
 *     Those accumulate in fp32.  BASELINE.json asks for fp64 accumulation, so the main entry points
 *     accumulate in double from the caller's (already dtype-rounded) float inputs; the *_f32 variants
 *     follow the Swift oracles operation-for-operation in float and exist to pin the double version
 *     to the reference's own 2e-5 tolerance (MFA/Tests/FlashAttentionTests/Attention/SquareAttentionTest.swift:557-571).
 *   - mask rules of the Metal kernel (no numeric test pins them upstream, kernel source is the definition)
 *       MFA/Sources/FlashAttention/Attention/AttentionKernel/AttentionKernel+Softmax.swift:445   causal: masked iff col > row
 *       ...AttentionKernel+Softmax.swift:450                                                       window: masked iff row > col + W
 *       ...AttentionKernel+Softmax.swift:316-329                                                   external additive mask [B,H,Sq,Skv]
 *   - side-output conventions (...AttentionKernel+Caching.swift:396-400, ...Softmax.swift:230-233)
 *       L = log2(e) * logsumexp(scale * S)          D = scale * sum_d dO*O
 *   - quantiser contract
 *       MFA/Sources/FlashAttention/GEMM/GEMMQuantization.swift:305-350,424-479   scale = absmax/127 (int8) | absmax/7 (int4)
 *       ...GEMMQuantization.swift:487-521                                         q = clamp(round(x/scale)), int4 nibble packing
 *       ...GEMMQuantization.swift:567-623                                         2-D block-wise, row-major block index
 *       MFA/Sources/FlashAttention/GEMM/GEMMRuntimeQuantization.swift:80-181      GPU variant: scale = max(absmax/den, 1e-8)
 *       ...GEMMHeaders.swift:694-695,757-772                                      dequant (q - zp) * scale, low nibble first
 *
 * Parity pinning: the quantiser is pinned by the reference's known-answer tests
 * (QuantizedAttentionTest.swift:30-59,61-161; Tests/QuantizationTests/QuantizationTests.swift:7-128).
 * The reference stores no attention output tensors; attention parity is pinned against golden vectors
 * produced by the reference's own Python test oracle (examples/pytorch-custom-op-ffi/tests/conftest.py:165-181,
 * torch SDPA on CPU) -- see tests/golden/make_golden.py.
 *
 * Layout everywhere: row-major BHSD, Q [B,H,Sq,D], K/V [B,H,Skv,D] (the layout the kernel really uses,
 * AttentionKernel+Source.swift:104-121), O like Q, L/D [B,H,Sq].
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LOG2E 1.4426950408889634

/* mask_mode: how an external additive mask enters the logits.
 *   0 = PyTorch semantics  softmax(scale*QK^T + mask)   (the B200 contract)
 *   1 = reference kernel   softmax(scale*(QK^T + mask)) (AttentionKernel+Softmax.swift:328-329 then :658,:791) */
typedef struct {
  int B, H, Sq, Skv, D;
  float scale;
  int causal;      /* 0/1 */
  int window;      /* <0: none; else masked iff row > col + window */
  int mask_mode;   /* see above */
} oracle_dims_t;

static inline int is_masked(const oracle_dims_t* p, int row, int col) {
  if (p->causal && col > row) return 1;
  if (p->window >= 0 && row > col + p->window) return 1;
  return 0;
}

/* logits for one row, in double: z[c] = scale*s + mask (or scale*(s+mask)); masked -> -inf */
static void logits_row(const oracle_dims_t* p, const float* q_row, const float* k_bh,
                       const float* mask_row, int row, double* z) {
  for (int c = 0; c < p->Skv; ++c) {
    if (is_masked(p, row, c)) { z[c] = -INFINITY; continue; }
    const float* kr = k_bh + (size_t)c * p->D;
    double s = 0.0;
    for (int d = 0; d < p->D; ++d) s += (double)q_row[d] * (double)kr[d];
    if (mask_row) {
      double m = (double)mask_row[c];
      z[c] = (p->mask_mode == 1) ? (double)p->scale * (s + m) : (double)p->scale * s + m;
    } else {
      z[c] = (double)p->scale * s;
    }
  }
}

/* softmax statistics of a row: returns natural-log LSE, fills prob[] (all zeros and -inf LSE if fully masked) */
static double softmax_row(int n, const double* z, double* prob) {
  double mx = -INFINITY;
  for (int c = 0; c < n; ++c) if (z[c] > mx) mx = z[c];
  if (mx == -INFINITY) { for (int c = 0; c < n; ++c) prob[c] = 0.0; return -INFINITY; }
  double sum = 0.0;
  for (int c = 0; c < n; ++c) sum += exp(z[c] - mx);
  double lse = mx + log(sum);
  for (int c = 0; c < n; ++c) prob[c] = exp(z[c] - lse);
  return lse;
}

/* Forward: O = softmax(logits) V, L = log2e * LSE.  mask: NULL or dense additive fp32 [B,H,Sq,Skv].
 * o and/or lse may be NULL.  Returns 0. */
int oracle_attention_forward(const float* q, const float* k, const float* v, const float* mask,
                             float* o, float* lse, int B, int H, int Sq, int Skv, int D,
                             float scale, int causal, int window, int mask_mode) {
  oracle_dims_t p = {B, H, Sq, Skv, D, scale, causal, window, mask_mode};
  long rows = (long)B * H * Sq;
#pragma omp parallel
  {
    double* z = (double*)malloc(sizeof(double) * (size_t)(Skv > 0 ? Skv : 1) * 2);
    double* pr = z + (Skv > 0 ? Skv : 1);
    double* acc = (double*)malloc(sizeof(double) * (size_t)(D > 0 ? D : 1));
#pragma omp for schedule(dynamic, 16)
    for (long r = 0; r < rows; ++r) {
      long bh = r / Sq; int row = (int)(r % Sq);
      const float* qr = q + (size_t)r * D;
      const float* kb = k + (size_t)bh * Skv * D;
      const float* vb = v + (size_t)bh * Skv * D;
      const float* mr = mask ? mask + (size_t)r * Skv : NULL;
      logits_row(&p, qr, kb, mr, row, z);
      double l = softmax_row(Skv, z, pr);
      if (lse) lse[r] = (float)(l * LOG2E);
      if (o) {
        for (int d = 0; d < D; ++d) acc[d] = 0.0;
        for (int c = 0; c < Skv; ++c) {
          double pc = pr[c];
          if (pc == 0.0) continue;
          const float* vr = vb + (size_t)c * D;
          for (int d = 0; d < D; ++d) acc[d] += pc * (double)vr[d];
        }
        for (int d = 0; d < D; ++d) o[(size_t)r * D + d] = (float)acc[d];
      }
    }
    free(z); free(acc);
  }
  return 0;
}

/* Backward (Network.swift:206-409 / cpuReferenceBackward): given dO recompute P, then
 *   Dterm = sum_d dO*O (stored as scale*Dterm, reference convention A3), dP = dO V^T,
 *   dS = P*(dP - Dterm)*scale, dQ = dS K, dK = dS^T Q, dV = P^T dO.   Any output may be NULL. */
int oracle_attention_backward(const float* q, const float* k, const float* v, const float* mask,
                              const float* d_o, float* dq, float* dk, float* dv, float* dterm,
                              int B, int H, int Sq, int Skv, int D,
                              float scale, int causal, int window, int mask_mode) {
  oracle_dims_t p = {B, H, Sq, Skv, D, scale, causal, window, mask_mode};
  long BH = (long)B * H;
#pragma omp parallel for schedule(dynamic, 1)
  for (long bh = 0; bh < BH; ++bh) {
    size_t kvn = (size_t)Skv * D;
    double* z = (double*)malloc(sizeof(double) * (size_t)(Skv + 1) * 2);
    double* pr = z + Skv + 1;
    double* orow = (double*)malloc(sizeof(double) * (size_t)(D + 1) * 2);
    double* dqrow = orow + D + 1;
    double* dka = dk ? (double*)calloc(kvn + 1, sizeof(double)) : NULL;
    double* dva = dv ? (double*)calloc(kvn + 1, sizeof(double)) : NULL;
    const float* kb = k + (size_t)bh * kvn;
    const float* vb = v + (size_t)bh * kvn;
    for (int row = 0; row < Sq; ++row) {
      size_t r = (size_t)bh * Sq + row;
      const float* qr = q + r * D;
      const float* dor = d_o + r * D;
      const float* mr = mask ? mask + r * Skv : NULL;
      logits_row(&p, qr, kb, mr, row, z);
      softmax_row(Skv, z, pr);
      for (int d = 0; d < D; ++d) { orow[d] = 0.0; dqrow[d] = 0.0; }
      for (int c = 0; c < Skv; ++c) {
        if (pr[c] == 0.0) continue;
        const float* vr = vb + (size_t)c * D;
        for (int d = 0; d < D; ++d) orow[d] += pr[c] * (double)vr[d];
      }
      double dt = 0.0;
      for (int d = 0; d < D; ++d) dt += orow[d] * (double)dor[d];
      if (dterm) dterm[r] = (float)(dt * (double)scale);
      for (int c = 0; c < Skv; ++c) {
        double pc = pr[c];
        if (pc == 0.0) continue;
        const float* vr = vb + (size_t)c * D;
        const float* kr = kb + (size_t)c * D;
        double dp = 0.0;
        for (int d = 0; d < D; ++d) dp += (double)dor[d] * (double)vr[d];
        double ds = pc * (dp - dt) * (double)scale;
        for (int d = 0; d < D; ++d) dqrow[d] += ds * (double)kr[d];
        if (dka) for (int d = 0; d < D; ++d) dka[(size_t)c * D + d] += ds * (double)qr[d];
        if (dva) for (int d = 0; d < D; ++d) dva[(size_t)c * D + d] += pc * (double)dor[d];
      }
      if (dq) for (int d = 0; d < D; ++d) dq[r * D + d] = (float)dqrow[d];
    }
    if (dk) for (size_t i = 0; i < kvn; ++i) dk[(size_t)bh * kvn + i] = (float)dka[i];
    if (dv) for (size_t i = 0; i < kvn; ++i) dv[(size_t)bh * kvn + i] = (float)dva[i];
    free(z); free(orow); free(dka); free(dva);
  }
  return 0;
}

/* Single-head fp32 restatement of the Swift oracle (Network.swift:136-190,298-321): same operation order,
 * float accumulators, expf/logf.  Used only to pin the double version against the reference's arithmetic. */
int oracle_attention_forward_f32(const float* q, const float* k, const float* v, float* o, float* lse_nat,
                                 int Sq, int Skv, int D) {
  float scale = 1.0f / sqrtf((float)D);
  float* s = (float*)malloc(sizeof(float) * (size_t)(Skv + 1));
  for (int r = 0; r < Sq; ++r) {
    for (int c = 0; c < Skv; ++c) {
      float dot = 0.f;
      for (int d = 0; d < D; ++d) dot += q[(size_t)r * D + d] * k[(size_t)c * D + d];
      s[c] = dot;
    }
    float mx = -FLT_MAX;
    for (int c = 0; c < Skv; ++c) mx = fmaxf(mx, scale * s[c]);
    float sum = 0.f;
    for (int c = 0; c < Skv; ++c) sum += expf(scale * s[c] - mx);
    float l = mx + logf(sum);
    if (lse_nat) lse_nat[r] = l;
    for (int c = 0; c < Skv; ++c) s[c] = expf(scale * s[c] - l);
    for (int d = 0; d < D; ++d) {
      float acc = 0.f;
      for (int c = 0; c < Skv; ++c) acc += s[c] * v[(size_t)c * D + d];
      o[(size_t)r * D + d] = acc;
    }
  }
  free(s);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Quantiser.  x is a row-major [rows, cols] fp32 view (values already rounded to the source dtype).
 * Blocks are block_rows x block_cols tiles, scales laid out row-major over blocks
 * (GEMMQuantization.swift:567-584).  Granularities map as
 *   tensor-wise : block_rows = rows, block_cols = cols   (GEMMQuantization.swift:305-350)
 *   row-wise    : block_rows = 1,    block_cols = cols   (:424-479)
 *   2-D block   : block_rows = block_cols = bs           (:567-623)
 *   token-block : block_rows = 64,   block_cols = D      (B200 contract)
 * bits = 8 -> int8 codes (rows*cols bytes); bits = 4 -> packed nibbles, element 2i in the low nibble,
 *   stored value = clamp(q,-8,7)+8, (rows*cols+1)/2 bytes, padding nibble = 8 (i.e. q=0) (:501-516,
 *   GEMMRuntimeQuantization.swift:128-140).
 * clamp_scale_min: 0 -> scale = absmax/den exactly (CPU path); >0 -> scale = max(absmax/den, v)
 *   (GPU path uses 1e-8f, GEMMRuntimeQuantization.swift:89,459).  A zero scale quantises to code 0.
 * Rounding: x / scale in fp32 then roundf (half away from zero) == Swift round()/MSL round().
 * ------------------------------------------------------------------------------------------- */
static inline int quant_code(float x, float scale, int bits) {
  if (!(scale > 0.f)) return 0;
  float r = roundf(x / scale);
  int lo = bits == 8 ? -128 : -8, hi = bits == 8 ? 127 : 7;
  if (r < (float)lo) return lo;
  if (r > (float)hi) return hi;
  return (int)r;
}

long oracle_quant_num_blocks(long rows, long cols, long block_rows, long block_cols) {
  long nbr = (rows + block_rows - 1) / block_rows, nbc = (cols + block_cols - 1) / block_cols;
  return nbr * nbc;
}

int oracle_quantize(const float* x, long rows, long cols, long block_rows, long block_cols, int bits,
                    float clamp_scale_min, uint8_t* codes, float* scales) {
  if (bits != 8 && bits != 4) return 1;
  if (block_rows <= 0 || block_cols <= 0) return 1;
  long nbr = (rows + block_rows - 1) / block_rows, nbc = (cols + block_cols - 1) / block_cols;
  float den = bits == 8 ? 127.0f : 7.0f;
#pragma omp parallel for schedule(static)
  for (long b = 0; b < nbr * nbc; ++b) {
    long br = b / nbc, bc = b % nbc;
    long r0 = br * block_rows, r1 = r0 + block_rows < rows ? r0 + block_rows : rows;
    long c0 = bc * block_cols, c1 = c0 + block_cols < cols ? c0 + block_cols : cols;
    float amax = 0.f;
    for (long r = r0; r < r1; ++r)
      for (long c = c0; c < c1; ++c) { float a = fabsf(x[r * cols + c]); if (a > amax) amax = a; }
    float sc = amax / den;
    if (clamp_scale_min > 0.f && !(sc > clamp_scale_min)) sc = clamp_scale_min;
    scales[b] = sc;
  }
  long n = rows * cols;
  if (bits == 8) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
      long r = i / cols, c = i % cols;
      float sc = scales[(r / block_rows) * nbc + c / block_cols];
      ((int8_t*)codes)[i] = (int8_t)quant_code(x[i], sc, 8);
    }
  } else {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i += 2) {
      long r = i / cols, c = i % cols;
      int q0 = quant_code(x[i], scales[(r / block_rows) * nbc + c / block_cols], 4) + 8;
      int q1 = 8;
      if (i + 1 < n) {
        long r1 = (i + 1) / cols, c1 = (i + 1) % cols;
        q1 = quant_code(x[i + 1], scales[(r1 / block_rows) * nbc + c1 / block_cols], 4) + 8;
      }
      codes[i / 2] = (uint8_t)(q0 | (q1 << 4));
    }
  }
  return 0;
}

int oracle_dequantize(const uint8_t* codes, const float* scales, long rows, long cols, long block_rows,
                      long block_cols, int bits, float* out) {
  if (bits != 8 && bits != 4) return 1;
  long nbc = (cols + block_cols - 1) / block_cols;
  long n = rows * cols;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    long r = i / cols, c = i % cols;
    float sc = scales[(r / block_rows) * nbc + c / block_cols];
    int qv;
    if (bits == 8) qv = ((const int8_t*)codes)[i];
    else { uint8_t byte = codes[i / 2]; qv = (int)((i & 1) ? (byte >> 4) : (byte & 0xF)) - 8; }
    out[i] = (float)qv * sc;
  }
  return 0;
}

/* bf16 / fp16 rounding helpers so tests build inputs exactly as the reference does
 * (round-to-nearest-even bf16: Tests/MFAFFITests/MultiHeadFFITests.swift:626-635). */
void oracle_round_bf16(const float* in, float* out, uint16_t* bits_out, long n) {
  for (long i = 0; i < n; ++i) {
    uint32_t u; memcpy(&u, &in[i], 4);
    uint32_t r;
    if ((u & 0x7fffffffu) > 0x7f800000u) r = (u >> 16) | 0x40u;           /* NaN */
    else r = (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
    if (bits_out) bits_out[i] = (uint16_t)r;
    if (out) { uint32_t w = r << 16; memcpy(&out[i], &w, 4); }
  }
}

/* Interleaved-pair rotary embedding (Sources/MFABridge/MFABridge.swift:269-319, ROPE_KERNEL): for every pair i of a row,
 *   out[2i] = x[2i] c - x[2i+1] s,  out[2i+1] = x[2i+1] c + x[2i] s,  c / s = table[b * table_batch_stride + s * D + 2i]
 * (pair-duplicated fp32 tables, only the even entry is read; table_batch_stride = 0 or S * D); negate_sin = the inverse
 * rotation = the backward transform of dQ / dK.  src / dst contiguous [B, H, S, D] fp32. */
void oracle_rope_rotate(const float* src, float* dst, const float* cos_t, const float* sin_t, long B, long H, long S, long D,
                        long table_batch_stride, int negate_sin) {
  for (long b = 0; b < B; ++b)
    for (long h = 0; h < H; ++h)
      for (long s = 0; s < S; ++s)
        for (long pair = 0; pair < D / 2; ++pair) {
          const long base = (((b * H + h) * S + s) * D) + 2 * pair;
          const long t = b * table_batch_stride + s * D + 2 * pair;
          const float c = cos_t[t];
          float sn = sin_t[t];
          if (negate_sin) sn = -sn;
          const float x0 = src[base], x1 = src[base + 1];
          dst[base] = x0 * c - x1 * sn;
          dst[base + 1] = x1 * c + x0 * sn;
        }
}

/* bench.py --impl reference: torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm asks for all host
 * cores explicitly so the CPU baseline is not a one-thread strawman. */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  extern void omp_set_num_threads(int);
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
