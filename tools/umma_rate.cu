// Stand-alone tcgen05.mma issue-rate probe (sm_100a): how many cycles does one kind::i8 / kind::f16 MMA of shape
// 128 x N x (32 B of K) take when issued back to back by one thread, with A in tensor memory or in shared memory?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_build/umma_rate tools/umma_rate.cu
// Operand contents are irrelevant (zeros); no loads run during the timed region.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../omniserve_b200/csrc/ptx.cuh"

using namespace ob;

// kind::f16 with bf16 operands, f32 accumulate: c_format=1 [4,6), a_format=1 (bf16) [7,10), b_format=1 [10,13)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
OB_DEVICE void umma_f16_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
OB_DEVICE void umma_f16_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

template <int N, int KIND /*0 i8, 1 bf16*/, int A_TMEM>
__global__ void __launch_bounds__(128, 1) probe(long long* cycles, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 128 rows x 128 B, SW128
  uint8_t* sB = smem + 16384;         // 256 rows x 128 B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sA)), db = umma_desc_kmajor_sw128(smem_u32(sB));
      const uint32_t ta = tm + 256;
      constexpr uint32_t id = KIND == 0 ? umma_idesc_i8(128, N, true, true) : idesc_bf16(128, N);
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          if (KIND == 0) {
            if (A_TMEM) umma_i8_ts(tm, ta + a * 8, db + 2 * a, id, 1u);
            else umma_i8_ss(tm, da + 2 * a, db + 2 * a, id, 1u);
          } else {
            if (A_TMEM) umma_f16_ts(tm, ta + a * 8, db + 2 * a, id, 1u);
            else umma_f16_ss(tm, da + 2 * a, db + 2 * a, id, 1u);
          }
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

template <int N, int KIND, int A_TMEM>
static void run(const char* tag, int ctas) {
  long long* d;
  cudaMalloc(&d, 8);
  auto k = probe<N, KIND, A_TMEM>;
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  k<<<ctas, 128, smem>>>(d, iters);   // warm-up
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<ctas, 128, smem>>>(d, iters);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / (iters * 4.0);
  const double macs = 128.0 * N * (KIND == 0 ? 32 : 16);
  printf("%-34s ctas=%3d: %7.1f cycles/MMA  %7.0f MAC/clk/SM   kernel %.3f ms -> %.2f Pop/s chip  (%s)\n", tag, ctas, per,
         macs / per, ms, 2.0 * macs * iters * 4.0 * ctas / (ms * 1e-3) / 1e15, cudaGetErrorString(err));
  cudaFree(d);
}

int main() {
  for (int ctas : {1, 148}) {
    run<64, 0, 1>("i8  128x64x32   A=tmem", ctas);
    run<128, 0, 1>("i8  128x128x32  A=tmem", ctas);
    run<256, 0, 1>("i8  128x256x32  A=tmem", ctas);
    run<64, 0, 0>("i8  128x64x32   A=smem", ctas);
    run<128, 0, 0>("i8  128x128x32  A=smem", ctas);
    run<256, 0, 0>("i8  128x256x32  A=smem", ctas);
    run<128, 1, 0>("bf16 128x128x16 A=smem", ctas);
    run<256, 1, 0>("bf16 128x256x16 A=smem", ctas);
    run<128, 1, 1>("bf16 128x128x16 A=tmem", ctas);
  }
  return 0;
}
