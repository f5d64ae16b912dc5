// attn_fwd_tcq.cu -- quantised attention forward on the sm_100a tensor pipe (SageAttention2-style):
//     S_int = Q_i8 K_i8^T            tcgen05.mma kind::i8 (int32 accumulator in TMEM, 2x the bf16 MMA rate)
//     S     = S_int * qs[row block] * ks[key block]          (fp32, in the softmax warps)
//     P     = exp2(S c - m)  -> bf16 in TMEM,  O += (P * vs[key block]) V_codes      (kind::f16, V codes exact in bf16)
// Replaces the reference's quantised forward (metal-flash-attention/Sources/FlashAttention/Attention/
// QuantizedAttention.swift:71-91,358-463 -- attention on dequantised int8 operands with fp32 statistics) for symmetric
// int8 codes (zero_point 0) with per-tensor or per-block scales (block = a multiple of 64 tokens of one (b, h); the
// SageAttention2 contract SURVEY Q5 settles on), head_dim 128.  int4 operands are unpacked to int8 codes and V codes
// to bf16 by HBM-bound pre-passes below.  Kernel skeleton, pipeline and masking are those of attn_fwd_tc.cu.
//
// int32 -> fp32 without I2F (which shares the 16/clk MUFU pipe with ex2): fm = as_float(0x4B400000 + s) = 1.5*2^23 + s
// exactly (|s| <= 128*127*127 < 2^22), row max is taken on fm (monotone), and the exponent argument is one FFMA
//     x = fm * a_h - (1.5*2^23 * a_h + m),   a_h = qs * ks[h] * scale * log2(e)
// whose only rounding is that of the constant (~0.75 a_h absolute in log2 units, << int8 quantisation noise).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math_constants.h>

#include "common.h"
#include "sm100_ptx.cuh"
#include "tc_host.h"

namespace mfa {

namespace {

using namespace ptx;

constexpr int kThreads = 384;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.f;
constexpr float kMagic = 12582912.f;          // 1.5 * 2^23
constexpr int D = 128;
constexpr int QTILE = 128 * 128;              // int8 Q / K tile: 128 rows x 128 bytes (one swizzle chunk)
constexpr int VTILE = 128 * 128 * 2;          // bf16 V tile: two chunks
constexpr int CHB = 128 * 128;
constexpr int NS = 5;                         // ring stages (32 KB each; K tiles use the first half)
constexpr int kSmem = 2 * QTILE + NS * VTILE + 64 + 16 * NS + 16 + 1024;

struct FwdQParams {
  CUtensorMap tq, tk, tv;
  void* o;
  long long o_sb, o_sh, o_ss;
  float* lse;
  int o_dtype;
  int H, Hkv, Sq, Skv;
  float c;
  int causal, window;
  const float* qs; const float* ks; const float* vs;   // per-block scale arrays (device) or nullptr
  float qs1, ks1, vs1;                                  // per-tensor scales when the array is null
  int qbr, kbr, vbr;                                    // tokens per block (multiple of 64 for K / V)
  int nbq, nbk, nbv;                                    // blocks per (b, h)
  int sq_, sk_, sv_;                                    // scale-array stride per (b, h): nb, or 0 for a single device scale
};

__global__ void __launch_bounds__(kThreads, 1) fwd_tcq_kernel(const __grid_constant__ FwdQParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sQ = base, sKV = base + 2 * QTILE, sBar = sKV + NS * VTILE;
  auto q_full = [&](int t) { return sBar + 8 * t; };
  auto s_full = [&](int t) { return sBar + 16 + 8 * t; };
  auto p_full = [&](int t) { return sBar + 32 + 8 * t; };
  auto o_full = [&](int t) { return sBar + 48 + 8 * t; };
  auto kv_full = [&](int s) { return sBar + 64 + 8 * s; };
  auto kv_empty = [&](int s) { return sBar + 64 + 8 * NS + 8 * s; };
  const uint32_t tmem_slot = sBar + 64 + 16 * NS;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int qblk = p.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.H / p.Hkv);
  const int r0 = qblk * 256;
  const int nt = (r0 + 128 < p.Sq) ? 2 : 1;
  int klo, khi;
  visible_key_range(p.causal, p.window, p.Skv, r0, min(r0 + 256, p.Sq), klo, khi);
  const int j_lo = klo >> 7;
  const int n = khi > klo ? ((khi + 127) >> 7) - j_lo : 0;

  if (threadIdx.x == 256) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(q_full(t), 1); mbar_init(s_full(t), 1); mbar_init(p_full(t), 4); mbar_init(o_full(t), 1);
    }
    for (int s = 0; s < NS; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    fence_mbar_init();
  }
  if (warp == 9) {
    if (lane == 0) { prefetch_tmap(&p.tq); prefetch_tmap(&p.tk); prefetch_tmap(&p.tv); }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw)), 0);

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0 && n > 0) {
      mbar_arrive_expect_tx(q_full(0), QTILE);
      tma_load_4d(sQ, &p.tq, q_full(0), 0, r0, h, b);
      for (int it = 0; it < n; ++it) {
        const int row = (j_lo + it) * 128;
        int idx = 2 * it, s = idx % NS;
        mbar_wait(kv_empty(s), ((idx / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(kv_full(s), QTILE);
        tma_load_4d(sKV + s * VTILE, &p.tk, kv_full(s), 0, row, hk, b);
        if (it == 0 && nt == 2) {
          mbar_arrive_expect_tx(q_full(1), QTILE);
          tma_load_4d(sQ + QTILE, &p.tq, q_full(1), 0, r0 + 128, h, b);
        }
        idx = 2 * it + 1; s = idx % NS;
        mbar_wait(kv_empty(s), ((idx / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(kv_full(s), VTILE);
        tma_load_4d(sKV + s * VTILE, &p.tv, kv_full(s), 0, row, hk, b);
        tma_load_4d(sKV + s * VTILE + CHB, &p.tv, kv_full(s), 64, row, hk, b);
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    if (n > 0) {
      constexpr uint32_t IDESC_S = make_idesc(2, 1, 1, 0, 0, 128, 128);      // s32 += s8 * s8, K-major A and B
      constexpr uint32_t IDESC_O = make_idesc(1, 1, 1, 0, 1, 128, D);        // f32 += bf16 (tmem) * bf16, B MN-major
      const uint32_t q_lo = desc_lo(sQ, 16), k_lo = desc_lo(sKV, 16), v_lo = desc_lo(sKV, CHB);
      auto issue_s = [&](int t, int idx) {
        const uint32_t a0 = q_lo + t * (QTILE >> 4), b0 = k_lo + (idx % NS) * (VTILE >> 4);
#pragma unroll
        for (int kk = 0; kk < D / 32; ++kk)        // 32 int8 per MMA = 32 bytes of the 128-byte row
          mma_i8_ss_u(tmem + t * 128, a0 + kk * 2, kDescHiSw128, b0 + kk * 2, kDescHiSw128, IDESC_S, kk > 0);
      };
      auto issue_o = [&](int t, int idx, bool acc) {
        const uint32_t b0 = v_lo + (idx % NS) * (VTILE >> 4);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          mma_f16_ts_u(tmem + 256 + t * D, tmem + t * 128 + kk * 8, b0 + kk * (2048 >> 4), kDescHiSw128, IDESC_O,
                       (acc || kk > 0) ? 1u : 0u);
      };
      auto wait_full = [&](int idx) { mbar_wait(kv_full(idx % NS), (idx / NS) & 1); };
      wait_full(0);
      for (int t = 0; t < nt; ++t) {
        mbar_wait(q_full(t), 0);
        tc_fence_after();
        issue_s(t, 0);
        tc_commit_u(s_full(t));
      }
      tc_commit_u(kv_empty(0));
      for (int it = 0; it < n; ++it) {
        const int vi = 2 * it + 1, ki = 2 * it + 2;
        wait_full(vi);
        for (int t = 0; t < nt; ++t) {
          mbar_wait(p_full(t), it & 1);
          tc_fence_after();
          issue_o(t, vi, it > 0);
          if (t == nt - 1) tc_commit_u(kv_empty(vi % NS));
          if (it + 1 < n) {
            if (t == 0) { wait_full(ki); tc_fence_after(); }
            issue_s(t, ki);
            tc_commit_u(s_full(t));
            if (t == nt - 1) tc_commit_u(kv_empty(ki % NS));
          } else {
            tc_commit_u(o_full(t));
          }
        }
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ softmax warpgroups
    const int t = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const int r = r0 + t * 128 + row;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_base + t * 128;
    const uint32_t tO = tmem + lane_base + 256 + t * D;
    float m = -CUDART_INF_F, l = 0.f;                     // m in scaled log2 units
    const int chi = p.causal ? min(p.Skv - 1, r) : p.Skv - 1;
    const int clo = p.window >= 0 ? max(0, r - p.window) : 0;
    const int rq = min(r, p.Sq - 1);
    const float qsc = (p.qs ? p.qs[((size_t)b * p.H + h) * p.sq_ + rq / p.qbr] : p.qs1) * p.c;
    const float* ksp = p.ks ? p.ks + ((size_t)b * p.Hkv + hk) * p.sk_ : nullptr;
    const float* vsp = p.vs ? p.vs + ((size_t)b * p.Hkv + hk) * p.sv_ : nullptr;
    const bool v_blocks = vsp != nullptr;

    if (t < nt) {
      for (int it = 0; it < n; ++it) {
        const int c0 = (j_lo + it) * 128;
        // scales of the two 64-key halves of this tile (blocks are multiples of 64 keys)
        float a0 = qsc * p.ks1, a1 = a0, v0 = 1.f, v1 = 1.f;
        if (ksp) {
          a0 = qsc * ksp[min(c0 / p.kbr, p.nbk - 1)];
          a1 = qsc * ksp[min((c0 + 64) / p.kbr, p.nbk - 1)];
        }
        if (v_blocks) {
          v0 = vsp[min(c0 / p.vbr, p.nbv - 1)];
          v1 = vsp[min((c0 + 64) / p.vbr, p.nbv - 1)];
        }
        mbar_wait(s_full(t), it & 1);
        tc_fence_after();
        uint32_t su[128];
        tmem_ld_x32(tS, su);
        tmem_ld_x32(tS + 32, su + 32);
        tmem_ld_x32(tS + 64, su + 64);
        tmem_ld_x32(tS + 96, su + 96);
        tmem_wait_ld();
        float* s = reinterpret_cast<float*>(su);
#pragma unroll
        for (int i = 0; i < 128; ++i) su[i] += 0x4B400000u;                    // fm = 1.5*2^23 + s_int, exact
        const bool need_mask = (c0 < clo) || (c0 + 127 > chi);
        const bool any_mask = __any_sync(0xffffffffu, need_mask);
        if (any_mask) {
          const int lo_i = clo - c0, hi_i = chi - c0;
#pragma unroll
          for (int i = 0; i < 128; ++i) s[i] = (i < lo_i || i > hi_i) ? -CUDART_INF_F : s[i];
        }
        float mxa = s[0], mxb = s[1], mxc = s[64], mxd = s[65];
#pragma unroll
        for (int i = 2; i < 64; i += 2) {
          mxa = fmaxf(mxa, s[i]); mxb = fmaxf(mxb, s[i + 1]); mxc = fmaxf(mxc, s[64 + i]); mxd = fmaxf(mxd, s[65 + i]);
        }
        // per-half maxima back to scaled log2 units; a_h > 0 so the max commutes with the scaling
        const float mh0 = (fmaxf(mxa, mxb) - kMagic) * a0, mh1 = (fmaxf(mxc, mxd) - kMagic) * a1;
        const float mx = fmaxf(mh0, mh1);                                        // -inf if every key is masked
        float m_new = fmaxf(m, mx);
        const bool grow = (m_new - m) > kRescaleThreshold;
        if (!grow) m_new = m;
        if (__any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? ex2(m - m_new) : 1.f;
          l *= alpha;
          if (it > 0) {
#pragma unroll
            for (int ch = 0; ch < D / 16; ++ch) {           // 16 columns at a time: the 128 scores stay live in registers
              uint32_t ou[16];
              tmem_ld_x16(tO + ch * 16, ou);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) ou[i] = __float_as_uint(__uint_as_float(ou[i]) * alpha);
              tmem_st_x16(tO + ch * 16, ou);
            }
          }
        }
        m = m_new;
        const float mm = (m == -CUDART_INF_F) ? 0.f : m;
        float nk0 = -fmaf(kMagic, a0, mm), nk1 = -fmaf(kMagic, a1, mm);
        // The folded constant costs ~0.75 a_h of absolute error in the exponent; with coarse scales (int4, or int8 data
        // of large magnitude) remove the offset exactly first (one more FADD per element, warp-uniform choice).
        if (__any_sync(0xffffffffu, fmaxf(a0, a1) > 0x1p-11f)) {
#pragma unroll
          for (int i = 0; i < 128; ++i) s[i] -= kMagic;
          nk0 = nk1 = -mm;
        }
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {                 // packed pairs overwrite the registers of scores already consumed
          const float ah = hh ? a1 : a0, nk = hh ? nk1 : nk0, vh = hh ? v1 : v0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float p0 = ex2(fmaf(s[64 * hh + 2 * i], ah, nk));
            const float p1 = ex2(fmaf(s[64 * hh + 2 * i + 1], ah, nk));
            sum0 += p0; sum1 += p1;
            su[32 * hh + i] = v_blocks ? pack_bf16(p0 * vh, p1 * vh) : pack_bf16(p0, p1);
          }
        }
        tmem_st_x32(tS, su);
        tmem_st_x32(tS + 32, su + 32);
        l += sum0 + sum1;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(t));
      }
      if (n > 0) {
        mbar_wait(o_full(t), 0);
        tc_fence_after();
      }
      const float inv = (l > 0.f ? 1.f / l : 0.f) * (v_blocks ? 1.f : p.vs1);
      const bool live = r < p.Sq;
      const size_t orow = (size_t)b * p.o_sb + (size_t)h * p.o_sh + (size_t)r * p.o_ss;
#pragma unroll
      for (int ch = 0; ch < D / 32; ++ch) {
        uint32_t ou[32];
        if (n > 0) {
          tmem_ld_x32(tO + ch * 32, ou);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) ou[i] = 0u;
        }
        if (live) {
          if (p.o_dtype == kF32) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.o) + orow + ch * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              dst[i] = make_float4(__uint_as_float(ou[4 * i]) * inv, __uint_as_float(ou[4 * i + 1]) * inv,
                                   __uint_as_float(ou[4 * i + 2]) * inv, __uint_as_float(ou[4 * i + 3]) * inv);
          } else {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.o) + orow + ch * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t wv[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float a = __uint_as_float(ou[8 * i + 2 * k]) * inv, bb = __uint_as_float(ou[8 * i + 2 * k + 1]) * inv;
                wv[k] = p.o_dtype == kBF16 ? pack_bf16(a, bb) : pack_f16(a, bb);
              }
              dst[i] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            }
          }
        }
      }
      if (live && p.lse) p.lse[((size_t)b * p.H + h) * p.Sq + r] = l > 0.f ? m + log2f(l) : -CUDART_INF_F;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

// ---- HBM-bound pre-passes -------------------------------------------------------------------------------------------
// int8 codes -> bf16 (exact: |code| <= 128 needs 8 significant bits); 16 codes per thread per step.
__global__ void codes_to_bf16_kernel(const int8_t* __restrict__ src, __nv_bfloat16* __restrict__ dst, uint64_t n16) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
    const int4 v = reinterpret_cast<const int4*>(src)[i];
    const int w[4] = {v.x, v.y, v.z, v.w};
    uint32_t out[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float f0 = (float)(int8_t)(w[k] & 0xff), f1 = (float)(int8_t)((w[k] >> 8) & 0xff);
      const float f2 = (float)(int8_t)((w[k] >> 16) & 0xff), f3 = (float)(int8_t)((w[k] >> 24) & 0xff);
      out[2 * k] = pack_bf16(f0, f1);
      out[2 * k + 1] = pack_bf16(f2, f3);
    }
    reinterpret_cast<uint4*>(dst)[2 * i] = make_uint4(out[0], out[1], out[2], out[3]);
    reinterpret_cast<uint4*>(dst)[2 * i + 1] = make_uint4(out[4], out[5], out[6], out[7]);
  }
}

// packed int4 (element 2i in the low nibble, stored +8: GEMMQuantization.swift:501-516) -> int8 codes; 32 codes / thread.
__global__ void int4_to_int8_kernel(const uint8_t* __restrict__ src, int8_t* __restrict__ dst, uint64_t n16) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(src)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t out[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // byte j of w[k] holds codes 2j (low nibble) and 2j+1 (high nibble)
      uint32_t lo = w[k] & 0x0f0f0f0fu, hi = (w[k] >> 4) & 0x0f0f0f0fu;
      // interleave: out bytes = lo0 hi0 lo1 hi1 | lo2 hi2 lo3 hi3, each minus 8
      const uint32_t e0 = __byte_perm(lo, hi, 0x5140), e1 = __byte_perm(lo, hi, 0x7362);
      out[2 * k] = __vsub4(e0, 0x08080808u);
      out[2 * k + 1] = __vsub4(e1, 0x08080808u);
    }
    reinterpret_cast<uint4*>(dst)[2 * i] = make_uint4(out[0], out[1], out[2], out[3]);
    reinterpret_cast<uint4*>(dst)[2 * i + 1] = make_uint4(out[4], out[5], out[6], out[7]);
  }
}

unsigned grid_for(uint64_t n, int threads) {
  uint64_t g = (n + threads - 1) / threads;
  const uint64_t cap = 148ull * 16;
  return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

bool scales_ok(const QuantView& q, bool need64) {
  if (q.zero_point != 0) return false;
  if (!q.scales || q.block_rows == 0) return q.block_rows == 0;     // one scale: by value, or scales[0] on the device
  if (q.block_rows < 0) return false;
  return !need64 || (q.block_rows % 64) == 0;
}

}  // namespace

cudaError_t launch_codes_to_bf16(const void* codes, void* dst, uint64_t n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  if (n % 16) return cudaErrorInvalidValue;
  codes_to_bf16_kernel<<<grid_for(n / 16, 256), 256, 0, st>>>(reinterpret_cast<const int8_t*>(codes),
                                                             reinterpret_cast<__nv_bfloat16*>(dst), n / 16);
  ++g_launch_count;
  return cudaGetLastError();
}

cudaError_t launch_int4_to_int8(const void* packed, void* codes, uint64_t n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  if (n % 32) return cudaErrorInvalidValue;
  int4_to_int8_kernel<<<grid_for(n / 32, 256), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(packed),
                                                            reinterpret_cast<int8_t*>(codes), n / 32);
  ++g_launch_count;
  return cudaGetLastError();
}

// Eligibility of the int8 tensor-core forward: int8 or int4 codes, symmetric, head_dim 128, contiguous operands, per-tensor
// scales or per-block scales with K / V blocks that are multiples of 64 tokens, no external mask.
bool fwd_tcq_eligible(const AttnParams& p) {
  if (getenv("MFA_DISABLE_TC") || getenv("MFA_DISABLE_TCQ")) return false;
  if (p.in_dtype != kI8 && p.in_dtype != kI4) return false;
  if (p.D != 128 || p.mask_kind != kMaskNone) return false;
  if (!(p.scale > 0.f) || p.Sq <= 0 || p.Skv <= 0 || p.B <= 0 || p.H <= 0 || p.Hkv <= 0 || p.H % p.Hkv) return false;
  if (p.B > 65535 || p.H > 65535) return false;
  if (!scales_ok(p.qq, false) || !scales_ok(p.qk, true) || !scales_ok(p.qv, true)) return false;
  const TensorView* vs[3] = {&p.q, &p.k, &p.v};
  const int64_t S[3] = {p.Sq, p.Skv, p.Skv}, Hn[3] = {p.H, p.Hkv, p.Hkv};
  for (int i = 0; i < 3; ++i) {      // contiguous BHSD (the pre-passes work on flat arrays)
    if (vs[i]->sd != 1 || vs[i]->ss != p.D || vs[i]->sh != S[i] * p.D || vs[i]->sb != Hn[i] * S[i] * p.D) return false;
    if (reinterpret_cast<uintptr_t>(vs[i]->ptr) & 15) return false;
  }
  if (p.o.sd != 1) return false;
  const int oes = dtype_bytes(p.o_dtype);
  if ((reinterpret_cast<uintptr_t>(p.o.ptr) & 15) || ((p.o.ss * oes) & 15) || ((p.o.sh * oes) & 15) || ((p.o.sb * oes) & 15))
    return false;
  return tc::encode_fn() != nullptr;
}

// scratch: at least  Skv*B*Hkv*D*2 (V as bf16)  [+ (Sq*B*H + Skv*B*Hkv)*D bytes of int8 codes when the input is int4].
size_t fwd_tcq_scratch_bytes(const AttnParams& p) {
  const size_t nq = (size_t)p.B * p.H * p.Sq * p.D, nkv = (size_t)p.B * p.Hkv * p.Skv * p.D;
  return nkv * 2 + (p.in_dtype == kI4 ? nq + 2 * nkv : 0) + 1024;
}

cudaError_t launch_fwd_tcq(const AttnParams& p, void* scratch, cudaStream_t st) {
  const size_t nq = (size_t)p.B * p.H * p.Sq * p.D, nkv = (size_t)p.B * p.Hkv * p.Skv * p.D;
  uint8_t* sc = reinterpret_cast<uint8_t*>(scratch);
  void* v16 = sc;
  const void *qc = p.q.ptr, *kc = p.k.ptr, *vc = p.v.ptr;
  cudaError_t e;
  if (p.in_dtype == kI4) {
    uint8_t* q8 = sc + ((nkv * 2 + 127) & ~(size_t)127);
    uint8_t* k8 = q8 + ((nq + 127) & ~(size_t)127);
    uint8_t* v8 = k8 + ((nkv + 127) & ~(size_t)127);
    if ((e = launch_int4_to_int8(qc, q8, nq, st)) != cudaSuccess) return e;
    if ((e = launch_int4_to_int8(kc, k8, nkv, st)) != cudaSuccess) return e;
    if ((e = launch_int4_to_int8(vc, v8, nkv, st)) != cudaSuccess) return e;
    qc = q8; kc = k8; vc = v8;
  }
  if ((e = launch_codes_to_bf16(vc, v16, nkv, st)) != cudaSuccess) return e;

  FwdQParams prm;
  TensorView tq = p.q, tk = p.k, tv = p.v;
  tq.ptr = qc; tk.ptr = kc; tv.ptr = v16;
  if (!tc::make_map(&prm.tq, tq, kI8, p.B, p.H, p.Sq, p.D) || !tc::make_map(&prm.tk, tk, kI8, p.B, p.Hkv, p.Skv, p.D) ||
      !tc::make_map(&prm.tv, tv, kBF16, p.B, p.Hkv, p.Skv, p.D))
    return cudaErrorInvalidValue;
  prm.o = const_cast<void*>(p.o.ptr);
  prm.o_sb = p.o.sb; prm.o_sh = p.o.sh; prm.o_ss = p.o.ss;
  prm.lse = p.lse; prm.o_dtype = p.o_dtype;
  prm.H = p.H; prm.Hkv = p.Hkv; prm.Sq = p.Sq; prm.Skv = p.Skv;
  prm.c = p.scale * kLog2e;
  prm.causal = p.causal; prm.window = p.window;
  auto setq = [](const QuantView& q, int S, const float*& arr, float& one, int& br, int& nb, int& stride) {
    arr = q.scales; one = q.scale;
    const bool blocks = q.scales && q.block_rows > 0;
    br = blocks ? q.block_rows : 1 << 30;
    nb = blocks ? (S + q.block_rows - 1) / q.block_rows : 1;
    stride = blocks ? nb : 0;
  };
  setq(p.qq, p.Sq, prm.qs, prm.qs1, prm.qbr, prm.nbq, prm.sq_);
  setq(p.qk, p.Skv, prm.ks, prm.ks1, prm.kbr, prm.nbk, prm.sk_);
  setq(p.qv, p.Skv, prm.vs, prm.vs1, prm.vbr, prm.nbv, prm.sv_);
  static bool attr_set = false;
  if (!attr_set) {
    if ((e = cudaFuncSetAttribute(fwd_tcq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)) != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((p.Sq + 255) / 256, p.H, p.B);
  fwd_tcq_kernel<<<grid, kThreads, kSmem, st>>>(prm);
  ++g_launch_count;
  g_last_kernel = p.in_dtype == kI4 ? "fwd_tcq_int4_d128" : "fwd_tcq_int8_d128";
  return cudaGetLastError();
}

}  // namespace mfa
