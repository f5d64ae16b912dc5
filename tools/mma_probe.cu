// mma_probe.cu -- measures tcgen05.mma issue-to-completion throughput on sm_100a for the operand configurations the
// attention kernels use (cycles per MMA instruction, one CTA per SM, operands = whatever is in smem / TMEM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../universal-metal-flash-attention_b200/csrc/sm100_ptx.cuh"
using namespace mfa::ptx;

// mode 0: SS  A K-major, B K-major, N=NN (bf16)      -> like S = Q K^T
// mode 1: TS  A tmem,    B MN-major, N=NN            -> like O += P V
// mode 2: SS  A K-major, B MN-major
// mode 3: alternate mode 0 and mode 1
// mode 4: i8 SS K-major (N=NN)
// mode 5: f8 TS B MN-major
template <int MODE, int NN>
__global__ void __launch_bounds__(192, 1) probe(long long* out, int iters, int ld_warps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 65536, sBar = base + 131072, slot = sBar + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(sBar, 1); mbar_init(sBar + 8, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t I_SS = make_idesc(1, 1, 1, 0, 0, 128, NN);
      constexpr uint32_t I_TS = make_idesc(1, 1, 1, 0, 1, 128, NN);
      constexpr uint32_t I_I8 = make_idesc(2, 1, 1, 0, 0, 128, NN);
      constexpr uint32_t I_F8 = make_idesc(1, 0, 0, 0, 1, 128, NN);
      constexpr uint32_t I_F8K = make_idesc(1, 0, 0, 0, 0, 128, NN);
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
          if (MODE == 0 || (MODE == 3 && (i & 1) == 0))
            mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + off, 16, 1024), I_SS, 1);
          else if (MODE == 1 || MODE == 3)
            mma_f16_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
          else if (MODE == 2)
            mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
          else if (MODE == 6)        // SS, accumulator alternates between two TMEM regions every 8 MMAs
            mma_f16_ss(tmem + (i & 1) * 128, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + off, 16, 1024), I_SS, 1);
          else if (MODE == 7)        // TS, accumulator and A alternate every 8 MMAs
            mma_f16_ts(tmem + 256 + (i & 1) * 128, tmem + (i & 1) * 128 + kk * 8, smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
          else if (MODE == 8) {      // SS K-major / SS B-MN-major alternate (same accumulator)
            if (i & 1) mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
            else mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + off, 16, 1024), I_SS, 1);
          } else if (MODE == 9) {    // SS / TS alternate, same accumulator, TS B K-major too (same idesc)
            if (i & 1) mma_f16_ts(tmem + 0, tmem + 128 + kk * 8, smem_desc_sw128(sB + off, 16, 1024), I_SS, 1);
            else mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + off, 16, 1024), I_SS, 1);
          } else if (MODE == 10) {   // SS / TS alternate every 16 MMAs
            if (i & 2) mma_f16_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
            else mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + off, 16, 1024), I_SS, 1);
          } else if (MODE == 11) {   // SS fresh accumulate (first MMA of each group overwrites) -> like S = Q K^T each tile
            mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + off, 16, 1024), I_SS, kk > 0);
          } else if (MODE == 12) {   // SS / TS alternate, SS overwrites (kk==0), separate accumulators: the forward pattern
            if (i & 1) mma_f16_ts(tmem + 256, tmem + 0 + kk * 8, smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
            else mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + off, 16, 1024), I_SS, kk > 0);
          } else if (MODE == 13) {   // forward pattern with a commit after every group
            if (i & 1) mma_f16_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
            else mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + off, 16, 1024), I_SS, kk > 0);
            if (kk == 7) tc_commit(sBar + 8);
          }
          else if (MODE == 14) {     // SS / TS alternate every 64 MMAs
            if (i & 8) mma_f16_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
            else mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + off, 16, 1024), I_SS, 1);
          } else if (MODE == 15) {   // TS with B K-major, pure
            mma_f16_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + off, 16, 1024), I_SS, 1);
          } else if (MODE == 16) {   // SS / TS alternate, TS reads its B from the A buffer region (disjoint smem)
            if (i & 1) mma_f16_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sA + 32768 + kk * 2048, 16384, 1024), I_TS, 1);
            else mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + off, 16, 1024), I_SS, 1);
          } else if (MODE == 17) {   // alternate per single MMA
            if (kk & 1) mma_f16_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
            else mma_f16_ss(tmem + 0, smem_desc_sw128(sA + off, 16, 1024), smem_desc_sw128(sB + off, 16, 1024), I_SS, 1);
          }
          else if (MODE == 18) {     // int8 forward pattern: 4 x i8 SS (S = Q K^T) then 8 x f16 TS (O += P V): kind switches
            if (i & 1) mma_f16_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
            else if (kk < 4) mma_i8_ss(tmem + 0, smem_desc_sw128(sA + kk * 32, 16, 1024), smem_desc_sw128(sB + kk * 32, 16, 1024), I_I8, kk > 0);
          } else if (MODE == 19) {   // 4 x i8 SS then 4 x f8 TS (K = 32 each)
            if (i & 1) { if (kk < 4) mma_f8_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + kk * 4096, 16384, 1024), I_F8, 1); }
            else if (kk < 4) mma_i8_ss(tmem + 0, smem_desc_sw128(sA + kk * 32, 16, 1024), smem_desc_sw128(sB + kk * 32, 16, 1024), I_I8, kk > 0);
          } else if (MODE == 20) {   // i8 SS / f16 TS alternate every MMA
            if (kk & 1) mma_f16_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
            else mma_i8_ss(tmem + 0, smem_desc_sw128(sA + (kk >> 1) * 32, 16, 1024), smem_desc_sw128(sB + (kk >> 1) * 32, 16, 1024), I_I8, 1);
          } else if (MODE == 21) {   // 4 x f8 SS (e4m3 Q K^T) then 4 x f8 TS: one kind throughout
            if (i & 1) { if (kk < 4) mma_f8_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + kk * 4096, 16384, 1024), I_F8, 1); }
            else if (kk < 4) mma_f8_ss(tmem + 0, smem_desc_sw128(sA + kk * 32, 16, 1024), smem_desc_sw128(sB + kk * 32, 16, 1024), I_F8K, kk > 0);
          } else if (MODE == 22) {   // 4 x i8 SS then 2 x f16 TS, 4 times (the 4-part P hand-off granularity)
            if (i & 1) mma_f16_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + kk * 2048, 16384, 1024), I_TS, 1);
            else if (kk < 4) mma_i8_ss(tmem + 128 * (kk & 1), smem_desc_sw128(sA + kk * 32, 16, 1024), smem_desc_sw128(sB + kk * 32, 16, 1024), I_I8, 1);
          }
          else if (MODE == 4)
            mma_i8_ss(tmem + 0, smem_desc_sw128(sA + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                      smem_desc_sw128(sB + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024), I_I8, 1);
          else if (MODE == 5)
            mma_f8_ts(tmem + 256, tmem + 128 + kk * 8, smem_desc_sw128(sB + kk * 4096, 16384, 1024), I_F8, 1);
        }
      }
      tc_commit(sBar);
      mbar_wait(sBar, 0);
      long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      stop = 1;
    }
  } else if (warp >= 2 && warp < 2 + ld_warps) {
    // concurrent TMEM traffic like the softmax warps: ld 32 cols, st 16 cols, repeat
    const uint32_t t = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384;
    uint32_t r[32];
    while (!stop) {
      tmem_ld_x32(t, r); tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] += 1;
      tmem_st_x16(t, r); tmem_wait_st();
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int MODE, int NN>
void run(const char* name, long long* d, int ld_warps) {
  const int iters = 2000, smem = 131072 + 1024 + 64;
  cudaFuncSetAttribute(probe<MODE, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<MODE, NN><<<148, 192, smem>>>(d, iters, ld_warps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (iters * 8.0);
  const int kdepth = (MODE == 4 || MODE == 5) ? 32 : 16;
  if (MODE >= 18) {
    const int ideal = MODE == 18 ? 768 : MODE == 20 ? 1024 : MODE == 22 ? 768 : 512;
    printf("%-60s %7.1f clk per (S, PV) pair   ideal %d  %s\n", name, (double)h / (iters / 2.0), ideal, e == cudaSuccess ? "" : cudaGetErrorString(e));
    return;
  }
  printf("%-44s N=%3d ld_warps=%d  %7.1f clk/MMA  (%5.1f%% of 8192 dense-bf16 FLOP/clk/SM)  %s\n", name, NN, ld_warps, per,
         100.0 * (2.0 * 128 * NN * kdepth / per) / 8192.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  run<0, 128>("SS  A,B K-major (S = Q K^T)", d, 0);
  run<0, 64>("SS  A,B K-major", d, 0);
  run<0, 256>("SS  A,B K-major", d, 0);
  run<1, 128>("TS  A tmem, B MN-major (O += P V)", d, 0);
  run<1, 64>("TS  A tmem, B MN-major", d, 0);
  run<2, 128>("SS  A K-major, B MN-major", d, 0);
  run<3, 128>("alternating SS / TS", d, 0);
  run<0, 128>("SS  + concurrent tcgen05.ld/st", d, 4);
  run<1, 128>("TS  + concurrent tcgen05.ld/st", d, 4);
  run<6, 128>("SS  accumulator alternates", d, 0);
  run<7, 128>("TS  accumulator + A alternate", d, 0);
  run<8, 128>("SS  B K-major / B MN-major alternate", d, 0);
  run<9, 128>("SS / TS alternate, same D and idesc", d, 0);
  run<10, 128>("SS / TS alternate every 16 MMAs", d, 0);
  run<11, 128>("SS  overwrite at kk==0", d, 0);
  run<12, 128>("SS(overwrite) / TS alternate, P aliases S", d, 0);
  run<13, 128>("SS / TS alternate + commit per group", d, 0);
  run<14, 128>("SS / TS alternate every 64 MMAs", d, 0);
  run<15, 128>("TS  B K-major pure", d, 0);
  run<16, 128>("SS / TS alternate, disjoint smem", d, 0);
  run<17, 128>("SS / TS alternate every MMA", d, 0);
  run<4, 128>("i8 SS K-major (K=32 per MMA)", d, 0);
  run<5, 128>("f8 TS B MN-major (K=32 per MMA)", d, 0);
  run<18, 128>("int8 fwd pattern: 4 i8 SS + 8 f16 TS", d, 0);
  run<19, 128>("4 i8 SS + 4 f8 TS", d, 0);
  run<20, 128>("i8 SS / f16 TS alternate every MMA (4+4 per iter)", d, 0);
  run<21, 128>("4 f8 SS + 4 f8 TS (single kind)", d, 0);
  run<22, 128>("4 i8 SS (2 accumulators) + 8 f16 TS", d, 0);
  return 0;
}
