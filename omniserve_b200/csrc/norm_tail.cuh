// add + "RMSNorm" + per-token INT8 quant (+sum) of ONE row by 128 threads that synchronise on a named barrier -- the
// body of small_ops.cu: rmsnorm_quant_kernel<FUSE_SUM, ADD = true> (reference: layernorm_kernels.cu:194-331 plus the
// residual add of llama_w4a8_unpad.py:425,437) restated so that it can run as the TAIL of another kernel (the W4A8 GEMM
// whose output is the `delta` of the add).  Element-to-thread assignment (vector idx = tid + 128 i), per-thread
// accumulation order and the reduction trees are those of the stand-alone kernel launched with 128 threads, so results
// are bit-identical to the two-kernel chain (tests/test_gpu_gemm.py::test_gemm_with_fused_add_norm_quant_tail).
#pragma once
#include "ptx.cuh"

namespace ob {

constexpr int TAIL_THREADS = 128;
constexpr int TAIL_NV = 4;   // 16-byte vectors per thread: rows of up to 4096 halves

OB_DEVICE void tail_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

OB_DEVICE float tail_warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
OB_DEVICE float tail_warp_max(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
  return v;
}
// small_ops.cu: block_reduce2 for 4 warps
template <bool MAX0>
OB_DEVICE void tail_reduce2(float& a, float& b, float* red /*[64] shared*/, int tid) {
  a = MAX0 ? tail_warp_max(a) : tail_warp_sum(a);
  b = tail_warp_sum(b);
  const int w = tid >> 5, l = tid & 31;
  tail_bar();
  if (l == 0) { red[w] = a; red[32 + w] = b; }
  tail_bar();
  float x = l < 4 ? red[l] : (MAX0 ? -3.0e38f : 0.f);
  float y = l < 4 ? red[32 + l] : 0.f;
  a = MAX0 ? tail_warp_max(x) : tail_warp_sum(x);
  b = tail_warp_sum(y);
}

union TailV8 {
  uint4 u;
  __half2 h2[4];
  __half h[8];
};

// hidden_out[row] = hidden_in[row] + delta[row] (fp16); q[row], scale[row], sum[row] = norm-quant of it.
// `delta_row` was written earlier in THIS kernel by other CTAs (made visible by the caller's grid barrier): it is read
// with ld.global.cg, never through the non-coherent path.
template <bool FUSE_SUM>
OB_DEVICE void add_norm_quant_row(int tid, const __half* hidden_in_row, const __half* delta_row, __half* hidden_out_row,
                                  const __half* gamma, int8_t* q_row, __half* scale_ptr, __half* sum_ptr, int H, float eps,
                                  float* red) {
  const int nvec = H >> 3;
  const uint4* src = reinterpret_cast<const uint4*>(hidden_in_row);
  const uint4* dsrc = reinterpret_cast<const uint4*>(delta_row);
  const uint4* gsrc = reinterpret_cast<const uint4*>(gamma);
  TailV8 v[TAIL_NV];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < TAIL_NV; ++i) {
    const int idx = tid + i * TAIL_THREADS;
    if (idx < nvec) {
      v[i].u = ld_nc_v4(src + idx);
      TailV8 dl;
      dl.u = __ldcg(dsrc + idx);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i].h2[j] = __hadd2(v[i].h2[j], dl.h2[j]);
      reinterpret_cast<uint4*>(hidden_out_row)[idx] = v[i].u;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float f = __half2float(v[i].h[j]);
        s1 += f;
        s2 += f * f;
      }
    }
  }
  tail_reduce2<false>(s1, s2, red, tid);
  const float mean = __fdividef(s1, (float)H);
  const float rstd = rsqrtf(__fdividef(s2, (float)H) + eps);
  float amax = 1.013279e-06f;  // (half)1e-6f, layernorm_kernels.cu:279
  __half hsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) hsum[j] = __float2half_rn(0.f);
  float nf[TAIL_NV][8];
#pragma unroll
  for (int i = 0; i < TAIL_NV; ++i) {
    const int idx = tid + i * TAIL_THREADS;
    if (idx < nvec) {
      TailV8 g;
      g.u = __ldg(gsrc + idx);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float f = ((__half2float(v[i].h[j]) - mean) * rstd) * __half2float(g.h[j]);
        nf[i][j] = f;
        const __half hv = __float2half_rn(f);
        amax = fmaxf(amax, fabsf(__half2float(hv)));
        if (FUSE_SUM) hsum[j] = __hadd(hsum[j], hv);
      }
    }
  }
  float ps = 0.f;
  if (FUSE_SUM) {
#pragma unroll
    for (int j = 0; j < 8; ++j) ps += __half2float(hsum[j]);
  }
  tail_reduce2<true>(amax, ps, red, tid);
  if (tid == 0) {
    *scale_ptr = __float2half_rn(__fdividef(amax, 127.0f));
    if (FUSE_SUM) *sum_ptr = __float2half_rn(ps);
  }
  const float qs = __fdividef(127.0f, amax);
  uint2* dst = reinterpret_cast<uint2*>(q_row);
#pragma unroll
  for (int i = 0; i < TAIL_NV; ++i) {
    const int idx = tid + i * TAIL_THREADS;
    if (idx < nvec) {
      uint32_t b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = (uint32_t)(uint8_t)f2i8_rni_sat(nf[i][j] * qs);
      uint2 r;
      r.x = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
      r.y = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
      dst[idx] = r;
    }
  }
  tail_bar();   // `red` is reused by the next row
}

}  // namespace ob
