#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/omniserve_b200.h"  // OB_ERR_* codes

namespace ob {

struct W4A8GemmArgs {
  const int8_t* in_feats;    // [M, K] int8 row-major
  const int8_t* qweight;     // [N, K/2] int8 in the reference tile layout
  const int8_t* s2_scales;   // per-group only: [K/128, N]
  const int8_t* s2_zeros;    // per-group only: [K/128, N]  (= -z*s2, two's complement)
  const __half* wscales;     // [N]
  const __half* ascales;     // [M]
  const __half* w_szs;       // per-channel only: [N]
  const __half* a_ssums;     // per-channel only: [M]
  __half* out_feats;         // [M, ldc]
  int M, N, K, ldc;
  int force_bn = 0;          // testing knobs: 0 = auto
  int force_mode = -1;       // -1 auto, 0 = data-parallel tiles, 1 = stream-K
  int force_ctas = 0;
  // optional fused tail (extension): hidden_out = hidden_in + out_feats (fp16), then rms_norm_general(_fuse_sum) of it
  const __half* tail_hidden_in = nullptr;   // [M, N]
  __half* tail_hidden_out = nullptr;        // [M, N]
  const __half* tail_gamma = nullptr;       // [N]
  int8_t* tail_q = nullptr;                 // [M, N]
  __half* tail_scale = nullptr;             // [M]
  __half* tail_sum = nullptr;               // [M] or null
  float tail_eps = 0.f;
};

int w4a8_gemm_run(const W4A8GemmArgs& a, bool per_group, cudaStream_t st);

}  // namespace ob
