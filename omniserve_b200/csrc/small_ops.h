#pragma once
#include <algorithm>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "w4a8_gemm.h"  // error codes

namespace ob {
int quant_run(const __half* in, int8_t* out, __half* scale, __half* sum /*nullable*/, int T, int H, cudaStream_t st);
// delta / hidden_out nullable: when given, x = in + delta is formed first (fp16) and stored to hidden_out
int rmsnorm_quant_run(const __half* in, const __half* delta, __half* hidden_out, const __half* gamma, int8_t* out,
                      __half* scale, __half* sum /*nullable*/, int T, int H, float eps, cudaStream_t st);
int rmsnorm_f16_run(const __half* in, const __half* delta /*nullable*/, const __half* gamma, __half* out, int T, int H,
                    float eps, cudaStream_t st);
int silu_and_mul_run(const __half* in, __half* out, int T, int d, cudaStream_t st);
int silu_mul_quant_run(const __half* in, int8_t* out, __half* scale, __half* sum /*nullable*/, int T, int d,
                       cudaStream_t st);
// tensor-parallel all-reduce over NVLink peer memory fused into add + norm (+ quant); see small_ops.cu
struct PeerArgs {
  const __half* bufs[8];     // every rank's partial-sum buffer [T, H] as mapped in this process (symmetric memory)
  uint32_t* flags[8];        // every rank's flag array [max_blocks][8] (uint32, zero-initialised once)
  uint32_t* epoch;           // local [max_blocks] (uint32, zero-initialised once)
  int world, rank, max_blocks;
};
int peer_rmsnorm_quant_run(const __half* in, const PeerArgs& peer, __half* hidden_out, const __half* gamma, int8_t* out,
                           __half* scale, __half* sum /*nullable*/, int T, int H, float eps, cudaStream_t st);
int peer_rmsnorm_f16_run(const __half* in, const PeerArgs& peer, const __half* gamma, __half* out, int T, int H, float eps,
                         cudaStream_t st);
int add_run(const __half* a, const __half* b, __half* out, size_t n, cudaStream_t st);
}  // namespace ob
