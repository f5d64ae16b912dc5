// f8_probe.cu -- numerics probe for the fp8 P V path: tcgen05.mma kind::f8f6f4 with A (e4m3, four per 32-bit column) read
// from TMEM and B (e4m3) read MN-major from a 128-byte-swizzled shared-memory tile -- the layout O += P V needs with V kept
// [keys][head_dim] as TMA delivers it.  Checks D = A B against a host reference for the candidate encodings:
//   variant 0: B MN-major  (smem rows = k, 128 bytes of n per row), A bytes little-endian in the column (k = 4c + byte)
//   variant 1: same, A bytes big-endian
//   variant 2: B K-major   (smem rows = n, 128 bytes of k per row) -- sanity check of the A packing alone
// and prints the byte order of cvt.rn.satfinite.e4m3x2.f32.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f8_probe f8_probe.cu && ./f8_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include "../universal-metal-flash-attention_b200/csrc/sm100_ptx.cuh"
using namespace mfa::ptx;

__global__ void __launch_bounds__(128, 1) probe(const uint8_t* __restrict__ A, const uint8_t* __restrict__ Bm, float* __restrict__ Dout,
                                                int variant, uint32_t* cvt_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sBp = smem_raw + (base - raw);
  const uint32_t sB = base, sBar = base + 16384, slot = sBar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
  if (threadIdx.x == 0) { mbar_init(sBar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));
  // B tile, 128 rows x 128 bytes, 16-byte units XOR-swizzled with (row & 7) -- what a SWIZZLE_128B TMA box would write.
  // variant 0/1: row = k, byte = n (Bm is [k][n]); variant 2: row = n, byte = k (transpose on the fly)
  for (int idx = threadIdx.x; idx < 128 * 128; idx += 128) {
    const int r = idx >> 7, b = idx & 127;
    const uint8_t val = variant == 2 ? Bm[b * 128 + r] : Bm[r * 128 + b];
    sBp[r * 128 + (((b >> 4) ^ (r & 7)) << 4) + (b & 15)] = val;
  }
  uint32_t a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const uint32_t v = A[row * 128 + 4 * c + b];
      w |= v << (8 * (variant == 1 ? 3 - b : b));
    }
    a[c] = w;
  }
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  tmem_st_x32(tmem + lane_base, a);
  tmem_wait_st();
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(1, 0, 0, 0, variant == 2 ? 0 : 1, 128, 128);      // f32 += e4m3 x e4m3
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t bd = variant == 2 ? smem_desc_sw128(sB + kk * 32, 16, 1024) : smem_desc_sw128(sB + kk * 4096, 16384, 1024);
      mma_f8_ts(tmem + 256, tmem + kk * 8, bd, idesc, kk > 0);
    }
    tc_commit(sBar);
  }
  mbar_wait(sBar, 0);
  tc_fence_after();
  uint32_t d[128];
#pragma unroll
  for (int i = 0; i < 4; ++i) tmem_ld_x32(tmem + lane_base + 256 + 32 * i, d + 32 * i);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 128; ++i) Dout[row * 128 + i] = __uint_as_float(d[i]);
  if (threadIdx.x == 0 && cvt_out) {
    unsigned short e;
    const float one = 1.0f, two = 2.0f;
    asm volatile("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(e) : "f"(one), "f"(two));
    cvt_out[0] = e;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
  (void)lane;
}

static uint8_t to_e4m3(float x) { return (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3); }
static float from_e4m3(uint8_t b) {
  const int s = b >> 7, e = (b >> 3) & 15, m = b & 7;
  const float v = e == 0 ? ldexpf((float)m, -9) : ldexpf(1.f + m / 8.f, e - 7);
  return s ? -v : v;
}

int main() {
  const int N = 128 * 128;
  uint8_t *hA = (uint8_t*)malloc(N), *hB = (uint8_t*)malloc(N);
  float *ref = (float*)malloc(N * 4), *hD = (float*)malloc(N * 4);
  srand(7);
  const float vals[] = {0.f, 0.5f, 1.f, 1.5f, 2.f, -1.f, -0.25f, 3.f, 0.125f, -2.5f};
  for (int i = 0; i < N; ++i) { hA[i] = to_e4m3(vals[rand() % 10]); hB[i] = to_e4m3(vals[rand() % 10]); }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      float s = 0.f;
      for (int k = 0; k < 128; ++k) s += from_e4m3(hA[m * 128 + k]) * from_e4m3(hB[k * 128 + n]);
      ref[m * 128 + n] = s;
    }
  uint8_t *dA, *dB; float* dD; uint32_t* dC;
  cudaMalloc(&dA, N); cudaMalloc(&dB, N); cudaMalloc(&dD, N * 4); cudaMalloc(&dC, 16);
  cudaMemcpy(dA, hA, N, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int variant = 0; variant < 3; ++variant) {
    cudaMemset(dD, 0, N * 4);
    probe<<<1, 128, 32768>>>(dA, dB, dD, variant, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < N; ++i) { const double d = fabs(hD[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
    printf("variant %d (%s): max abs err %.4g, mismatches %d / %d   D[0][0..3] = %g %g %g %g (ref %g %g %g %g)\n", variant,
           variant == 0 ? "B MN-major, A little-endian" : variant == 1 ? "B MN-major, A big-endian" : "B K-major, A little-endian",
           maxerr, bad, N, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
  }
  uint32_t hc = 0;
  cudaMemcpy(&hc, dC, 4, cudaMemcpyDeviceToHost);
  printf("cvt.rn.satfinite.e4m3x2.f32 d, 1.0, 2.0 -> 0x%04x  (e4m3 1.0 = 0x38, 2.0 = 0x40: first operand lands in the %s byte)\n", hc,
         (hc >> 8) == 0x38 ? "HIGH" : "LOW");
  return 0;
}
