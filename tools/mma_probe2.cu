// mma_probe2.cu -- per-group completion times of mixed SS / TS tcgen05.mma streams (which group pays for the mix?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe2 mma_probe2.cu && ./mma_probe2
// One thread issues groups of G MMAs, each followed by a commit to its own mbarrier; a monitor warp waits on the
// barriers in order and stamps clock64 at each completion.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../universal-metal-flash-attention_b200/csrc/sm100_ptx.cuh"
using namespace mfa::ptx;

constexpr int kGroups = 32;

// kind of group g: pattern 0: alternate SS / TS;  1: all SS;  2: all TS;  3: SS SS TS TS;  4: i8 SS / f16 TS;  5: i8 SS / f8 TS
// 6: SS, TS alternate per MMA inside every group
template <int PAT, int G>
__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 65536, sBar = base + 131072, slot = sBar + 8 * kGroups + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int g = 0; g < kGroups; ++g) mbar_init(sBar + 8 * g, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));
  __shared__ long long stamps[kGroups + 1];
  if (warp == 0 && lane == 0) {
    constexpr uint32_t I_SS = make_idesc(1, 1, 1, 0, 0, 128, 128);
    constexpr uint32_t I_TS = make_idesc(1, 1, 1, 0, 1, 128, 128);
    constexpr uint32_t I_I8 = make_idesc(2, 1, 1, 0, 0, 128, 128);
    constexpr uint32_t I_F8 = make_idesc(1, 0, 0, 0, 1, 128, 128);
    for (int g = 0; g < kGroups; ++g) {
#pragma unroll
      for (int k = 0; k < G; ++k) {
        const int kk = k & 7;
        const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
        bool ts;
        if (PAT == 0 || PAT == 4 || PAT == 5) ts = g & 1;
        else if (PAT == 1) ts = false;
        else if (PAT == 2) ts = true;
        else if (PAT == 3) ts = (g >> 1) & 1;
        else ts = k & 1;
        if (!ts) {
          if (PAT == 4 || PAT == 5) mma_i8_ss(tmem + 0, smem_desc_sw128(sA + (kk & 3) * 32, 16, 1024), smem_desc_sw128(sB + (kk & 3) * 32, 16, 1024), I_I8, 1);
          else mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + off, 16, 1024), I_SS, 1);
        } else {
          if (PAT == 5) mma_f8_ts(tmem + 256, tmem + 128 + (kk & 3) * 8, smem_desc_sw128(sB + (kk & 3) * 4096, 16384, 1024), I_F8, 1);
          else mma_f16_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
        }
      }
      tc_commit(sBar + 8 * g);
    }
  } else if (warp == 2 && lane == 0) {
    stamps[0] = clock64();
    for (int g = 0; g < kGroups; ++g) { mbar_wait(sBar + 8 * g, 0); stamps[g + 1] = clock64(); }
    if (blockIdx.x == 0) for (int g = 0; g <= kGroups; ++g) out[g] = stamps[g];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int PAT, int G>
void run(const char* name, long long* d) {
  const int smem = 131072 + 2048;
  cudaFuncSetAttribute(probe<PAT, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe<PAT, G><<<148, 128, smem>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[kGroups + 1];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-40s G=%2d total %6lld |", name, G, h[kGroups] - h[0]);
  for (int g = 8; g < 20; ++g) printf(" %4lld", h[g + 1] - h[g]);
  printf("  %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 8 * (kGroups + 1));
  run<1, 8>("all SS", d);
  run<2, 8>("all TS", d);
  run<0, 8>("SS / TS alternate groups", d);
  run<0, 4>("SS / TS alternate groups", d);
  run<0, 2>("SS / TS alternate groups", d);
  run<0, 1>("SS / TS alternate groups", d);
  run<0, 16>("SS / TS alternate groups", d);
  run<3, 8>("SS SS TS TS groups", d);
  run<6, 8>("SS/TS per MMA inside groups", d);
  run<4, 8>("i8 SS(x8) / f16 TS(x8) groups", d);
  run<4, 4>("i8 SS(x4) / f16 TS(x4) groups", d);
  run<5, 4>("i8 SS(x4) / f8 TS(x4) groups", d);
  return 0;
}
