// attn_fwd_tc.cu -- fused attention forward for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM.
//
// Replaces the reference's `attention` kernel, type = forward
// (metal-flash-attention/Sources/FlashAttention/Attention/AttentionKernel/AttentionKernel+Source.swift:372-416,
//  online softmax +Softmax.swift:641-702, finalisation +Caching.swift:396-400) for bf16/fp16 operands with
// head_dim 64 or 128.  Same math, re-derived for Blackwell:
//
//   one CTA = 2 query tiles of 128 rows (one per softmax warpgroup) sharing one K/V stream
//   warp 9  : TMA producer   Q tiles once, then K_j, V_j tiles through an NS-stage mbarrier ring (128B swizzle)
//   warp 8  : MMA issuer     S_t = Q_t K_j^T   (tcgen05.mma SS, fp32 accumulator in TMEM columns [t*128, +128))
//                            O_t += P_t V_j    (tcgen05.mma TS: P read straight from TMEM, V MN-major from smem)
//   warps 0-3 / 4-7 : softmax warpgroup of tile 0 / 1; thread i owns row i of its tile (tcgen05.ld 32x32b), so
//                     row max / row sum need no shuffles; P is written back over S as packed 16-bit pairs.
//
// The two tiles ping-pong on the tensor pipe: while warpgroup 0 runs exp2 on S_0, the pipe works on tile 1.
// O is only rescaled when the running row max grew by more than 2^8 (the stale max is kept otherwise; the final
// division by l absorbs it), so the O read-modify-write in TMEM is rare.
// KV tiles that are fully hidden by the causal / sliding-window rule are never loaded (loop bounds), tiles that
// are partly hidden get an element mask, tiles that are fully visible skip the mask code.
// Outputs follow the reference contract: O fp32 (or fp16/bf16 on request) and L = m + log2(l) in log2 units.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <mutex>

#include "common.h"
#include "sm100_ptx.cuh"
#include "tc_host.h"

namespace mfa {

namespace {

using namespace ptx;

constexpr int kThreads = 384;           // warpgroups: softmax 0, softmax 1, {MMA, TMA, 2 idle warps}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.f;     // log2 units

struct FwdTcParams {
  CUtensorMap tq, tk, tv;
  void* o;
  long long o_sb, o_sh, o_ss;
  float* lse;
  int o_dtype;
  int H, Hkv, Sq, Skv;
  float c;                 // softmax_scale * log2(e)
  int causal, window;
};

template <int D>
struct Cfg {
  static constexpr int kTile = 128 * D * 2;              // bytes of one 128-row operand tile
  static constexpr int kChunks = D / 64;                 // 128-byte swizzle chunks per row
  static constexpr int kChunkBytes = 128 * 128;          // one chunk of 128 rows
  static constexpr int kStages = D == 128 ? 5 : 10;
  static constexpr int kBarBytes = 64 + 16 * kStages + 16;
  static constexpr int kSmem = (2 + kStages) * kTile + kBarBytes + 1024;
};

// 2^x for a pair of x <= ~8 on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, max rel err 8.6e-5),
// in packed f32x2 arithmetic: relieves the MUFU pipe, which is co-critical with the tensor pipe at head_dim 128
// (16 ex2/clk/SM vs 8192 FLOP/clk/SM).
__device__ __forceinline__ f32x2 exp2_poly2(f32x2 x) {
  float x0, x1;
  unpack2(x, x0, x1);
  x = pack2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const f32x2 magic = pack2(12582912.f, 12582912.f);                  // 1.5 * 2^23: low mantissa bits = floor(x)
  const f32x2 xr = add2_rm(x, magic);
  const f32x2 f = sub2(x, sub2(xr, magic));                           // [0, 1)
  f32x2 pz = fma2(f, pack2(0.07706582f, 0.07706582f), pack2(0.22764632f, 0.22764632f));
  pz = fma2(pz, f, pack2(0.69511649f, 0.69511649f));
  pz = fma2(pz, f, pack2(1.0f, 1.0f));
  float p0, p1, r0, r1;
  unpack2(pz, p0, p1);
  unpack2(xr, r0, r1);
  return pack2(__int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23)),
               __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23)));
}

// P = exp2(s c - m c) for the 128 scores of one row, packed to 16-bit pairs; row sum of the fp32 values.
// NP of every 8 element pairs take the polynomial (compile-time pattern, so the loop body is branch-free).
template <bool BF16, int NP>
__device__ __forceinline__ void exp_phase(const float* s, float c, float neg_mc, uint32_t* pk, float& sum0, float& sum1) {
  constexpr int kSel[5] = {0x00, 0x08, 0x22, 0x52, 0xAA};
  const f32x2 c2 = pack2(c, c), nm2 = pack2(neg_mc, neg_mc);
  f32x2 acc_a = pack2(0.f, 0.f), acc_b = acc_a;
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    const f32x2 x = fma2(pack2(s[2 * i], s[2 * i + 1]), c2, nm2);
    f32x2 pp;
    float p0, p1;
    if ((kSel[NP] >> (i & 7)) & 1) {
      pp = exp2_poly2(x);
      unpack2(pp, p0, p1);
    } else {
      float x0, x1;
      unpack2(x, x0, x1);
      p0 = ex2(x0); p1 = ex2(x1);
      pp = pack2(p0, p1);
    }
    if (i & 1) acc_b = add2(acc_b, pp); else acc_a = add2(acc_a, pp);
    pk[i] = BF16 ? pack_bf16(p0, p1) : pack_f16(p0, p1);
  }
  unpack2(add2(acc_a, acc_b), sum0, sum1);
}

// POLY = n in 0..4: n of every 8 element pairs go through exp2_poly2 instead of MUFU.EX2 (only on tiles without masking,
// so masked elements always get an exact zero weight).
template <int D, bool BF16, int POLY>
__global__ void __launch_bounds__(kThreads, 1) fwd_tc_kernel(const __grid_constant__ FwdTcParams p) {
  using C = Cfg<D>;
  constexpr int TILE = C::kTile, NS = C::kStages, CHB = C::kChunkBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sQ = base, sKV = base + 2 * TILE, sBar = sKV + NS * TILE;
  auto q_full = [&](int t) { return sBar + 8 * t; };
  auto s_full = [&](int t) { return sBar + 16 + 8 * t; };
  auto p_full = [&](int t) { return sBar + 32 + 8 * t; };
  auto o_full = [&](int t) { return sBar + 48 + 8 * t; };
  auto kv_full = [&](int s) { return sBar + 64 + 8 * s; };
  auto kv_empty = [&](int s) { return sBar + 64 + 8 * NS + 8 * s; };
  const uint32_t tmem_slot = sBar + 64 + 16 * NS;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for ptxas
  const int qblk = p.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;   // heavy blocks first
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
    reg_dealloc<40>();
    if (lane == 0 && n > 0) {
      auto load_tile = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int row, int head) {
        mbar_arrive_expect_tx(bar, TILE);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) tma_load_4d(dst + c * CHB, m, bar, c * 64, row, head, b);
      };
      load_tile(sQ, &p.tq, q_full(0), r0, h);
      for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int kv = 0; kv < 2; ++kv) {
          const int idx = 2 * it + kv, s = idx % NS, ph = (idx / NS) & 1;
          mbar_wait(kv_empty(s), ph ^ 1);
          load_tile(sKV + s * TILE, kv ? &p.tv : &p.tk, kv_full(s), (j_lo + it) * 128, hk);
          if (it == 0 && kv == 0 && nt == 2) load_tile(sQ + TILE, &p.tq, q_full(1), r0 + 128, h);
        }
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    reg_dealloc<40>();
    if (n > 0) {
      constexpr uint32_t FMT = BF16 ? 1u : 0u;
      constexpr uint32_t IDESC_S = make_idesc(1, FMT, FMT, 0, 0, 128, 128);
      constexpr uint32_t IDESC_O = make_idesc(1, FMT, FMT, 0, 1, 128, D);
      const uint32_t q_lo = desc_lo(sQ, 16), k_lo = desc_lo(sKV, 16), v_lo = desc_lo(sKV, CHB);
      auto issue_s = [&](int t, int idx) {
        const uint32_t a0 = q_lo + t * (TILE >> 4), b0 = k_lo + (idx % NS) * (TILE >> 4);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = ((kk >> 2) * CHB + (kk & 3) * 32) >> 4;
          mma_f16_ss_u(tmem + t * 128, a0 + off, kDescHiSw128, b0 + off, kDescHiSw128, IDESC_S, kk > 0);
        }
      };
      auto issue_o = [&](int t, int idx, bool acc) {
        const uint32_t b0 = v_lo + (idx % NS) * (TILE >> 4);
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
    reg_alloc<232>();
    const int t = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const int r = r0 + t * 128 + row;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_base + t * 128;
    const uint32_t tO = tmem + lane_base + 256 + t * D;
    const float c = p.c;
    float m = -CUDART_INF_F, l = 0.f;
    const int chi = p.causal ? min(p.Skv - 1, r) : p.Skv - 1;
    const int clo = p.window >= 0 ? max(0, r - p.window) : 0;

    if (t < nt) {
      for (int it = 0; it < n; ++it) {
        const int c0 = (j_lo + it) * 128;
        mbar_wait(s_full(t), it & 1);
        tc_fence_after();
        uint32_t su[128];
        tmem_ld_x32(tS, su);
        tmem_ld_x32(tS + 32, su + 32);
        tmem_ld_x32(tS + 64, su + 64);
        tmem_ld_x32(tS + 96, su + 96);
        tmem_wait_ld();
        float* s = reinterpret_cast<float*>(su);
        const bool need_mask = (c0 < clo) || (c0 + 127 > chi);
        const bool any_mask = __any_sync(0xffffffffu, need_mask);
        if (any_mask) {
          const int lo_i = clo - c0, hi_i = chi - c0;
#pragma unroll
          for (int i = 0; i < 128; ++i) s[i] = (i < lo_i || i > hi_i) ? -CUDART_INF_F : s[i];
        }
        float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
        for (int i = 4; i < 128; i += 4) {
          mx0 = fmaxf(mx0, s[i]); mx1 = fmaxf(mx1, s[i + 1]); mx2 = fmaxf(mx2, s[i + 2]); mx3 = fmaxf(mx3, s[i + 3]);
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        float m_new = fmaxf(m, mx);
        const bool grow = (m_new - m) * c > kRescaleThreshold;      // false when both are -inf (NaN compare)
        if (!grow) m_new = m;
        if (__any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? ex2((m - m_new) * c) : 1.f;    // m = -inf -> 0
          l *= alpha;
          if (it > 0) {
#pragma unroll
            for (int ch = 0; ch < D / 32; ++ch) {
              uint32_t ou[32];
              tmem_ld_x32(tO + ch * 32, ou);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) ou[i] = __float_as_uint(__uint_as_float(ou[i]) * alpha);
              tmem_st_x32(tO + ch * 32, ou);
            }
          }
        }
        m = m_new;
        const float neg_mc = (m == -CUDART_INF_F) ? 0.f : -m * c;
        uint32_t pk[64];
        float sum0, sum1;
        if (POLY > 0 && !any_mask) exp_phase<BF16, POLY>(s, c, neg_mc, pk, sum0, sum1);
        else exp_phase<BF16, 0>(s, c, neg_mc, pk, sum0, sum1);
        l += sum0 + sum1;
        tmem_st_x32(tS, pk);
        tmem_st_x32(tS + 32, pk + 32);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(t));
      }
      // ---------------------------------------------------------------- epilogue: O / l, L = m + log2(l)
      if (n > 0) {
        mbar_wait(o_full(t), 0);
        tc_fence_after();
      }
      const float inv = l > 0.f ? 1.f / l : 0.f;
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
              uint32_t w[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float a = __uint_as_float(ou[8 * i + 2 * k]) * inv, bb = __uint_as_float(ou[8 * i + 2 * k + 1]) * inv;
                w[k] = p.o_dtype == kBF16 ? pack_bf16(a, bb) : pack_f16(a, bb);
              }
              dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      }
      if (live && p.lse) p.lse[((size_t)b * p.H + h) * p.Sq + r] = l > 0.f ? fmaf(m, c, log2f(l)) : -CUDART_INF_F;
    }
  }
  else {
    reg_dealloc<40>();      // idle warps of the third warpgroup (setmaxnreg is warpgroup-wide)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ host side
using tc::encode_fn;
using tc::make_map;

bool view_ok(const TensorView& t, int64_t, int64_t Hn, int64_t B) { return tc::view_ok(t, Hn, B); }

int poly_setting() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MFA_FWD_POLY"); v = e ? atoi(e) : 2; }
  return v;
}

template <int D, bool BF16, int POLY>
cudaError_t launch_k(const FwdTcParams& prm, dim3 grid, cudaStream_t st) {
  static bool attr_set = false;
  auto kern = fwd_tc_kernel<D, BF16, POLY>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<D>::kSmem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  kern<<<grid, kThreads, Cfg<D>::kSmem, st>>>(prm);
  return cudaGetLastError();
}

template <int D, bool BF16>
cudaError_t launch(const FwdTcParams& prm, dim3 grid, cudaStream_t st) {
  switch (poly_setting()) {
    case 1: return launch_k<D, BF16, 1>(prm, grid, st);
    case 2: return launch_k<D, BF16, 2>(prm, grid, st);
    case 3: return launch_k<D, BF16, 3>(prm, grid, st);
    case 4: return launch_k<D, BF16, 4>(prm, grid, st);
    default: return launch_k<D, BF16, 0>(prm, grid, st);
  }
}

}  // namespace

bool fwd_tc_eligible(const AttnParams& p) {
  if (getenv("MFA_DISABLE_TC")) return false;
  if (p.in_dtype != kBF16 && p.in_dtype != kF16) return false;
  if (p.D != 64 && p.D != 128) return false;
  if (p.mask_kind != kMaskNone) return false;
  if (!(p.scale > 0.f) || p.Sq <= 0 || p.Skv <= 0 || p.B <= 0 || p.H <= 0 || p.Hkv <= 0 || p.H % p.Hkv) return false;
  if (p.B > 65535 || p.H > 65535) return false;
  if (!view_ok(p.q, p.Sq, p.H, p.B) || !view_ok(p.k, p.Skv, p.Hkv, p.B) || !view_ok(p.v, p.Skv, p.Hkv, p.B)) return false;
  if (p.o.sd != 1) return false;
  const int oes = dtype_bytes(p.o_dtype);
  if ((reinterpret_cast<uintptr_t>(p.o.ptr) & 15) || ((p.o.ss * oes) & 15) || ((p.o.sh * oes) & 15) || ((p.o.sb * oes) & 15))
    return false;
  return encode_fn() != nullptr;
}

cudaError_t launch_fwd_tc(const AttnParams& p, cudaStream_t st) {
  FwdTcParams prm;
  if (!make_map(&prm.tq, p.q, p.in_dtype, p.B, p.H, p.Sq, p.D) || !make_map(&prm.tk, p.k, p.in_dtype, p.B, p.Hkv, p.Skv, p.D) ||
      !make_map(&prm.tv, p.v, p.in_dtype, p.B, p.Hkv, p.Skv, p.D))
    return cudaErrorInvalidValue;
  prm.o = const_cast<void*>(p.o.ptr);
  prm.o_sb = p.o.sb; prm.o_sh = p.o.sh; prm.o_ss = p.o.ss;
  prm.lse = p.lse;
  prm.o_dtype = p.o_dtype;
  prm.H = p.H; prm.Hkv = p.Hkv; prm.Sq = p.Sq; prm.Skv = p.Skv;
  prm.c = p.scale * kLog2e;
  prm.causal = p.causal; prm.window = p.window;
  dim3 grid((p.Sq + 255) / 256, p.H, p.B);
  cudaError_t e;
  const bool bf = p.in_dtype == kBF16;
  if (p.D == 128) {
    e = bf ? launch<128, true>(prm, grid, st) : launch<128, false>(prm, grid, st);
    g_last_kernel = bf ? "fwd_tc_bf16_d128" : "fwd_tc_fp16_d128";
  } else {
    e = bf ? launch<64, true>(prm, grid, st) : launch<64, false>(prm, grid, st);
    g_last_kernel = bf ? "fwd_tc_bf16_d64" : "fwd_tc_fp16_d64";
  }
  ++g_launch_count;
  return e;
}

}  // namespace mfa
