// HBM-bound fused small ops around the W4A8 GEMMs: per-token INT8 activation quant (+sum),
// "RMSNorm"+quant(+sum) with the reference's quirks, plain rms_norm, SiLU*mul (+ fused quant).
//
// Replaces: /root/reference/kernels/csrc/fused_kernels.cu:57-142, layernorm_kernels.cu:194-364,
//           activation_kernels.cu:10-30.   One CTA per token row, 16-byte vector loads, the row is
// kept in registers so HBM sees exactly one read of the input and one write of the output
// (the reference re-reads the row 2-3x).  Fast-math intrinsics are used where the reference's
// `--use_fast_math` build uses them so that the INT8 codes agree with the rebuilt reference.
#include "launch.h"
#include "ptx.cuh"
#include "small_ops.h"

namespace ob {

constexpr int MAXV = 8;  // 16-byte vectors cached per thread

OB_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
OB_DEVICE float warp_max(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
  return v;
}

// Block-wide reduction of two values at once (op0 = sum or max, op1 = sum); result broadcast.
template <bool MAX0>
OB_DEVICE void block_reduce2(float& a, float& b, float* red /*[64]*/) {
  a = MAX0 ? warp_max(a) : warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) { red[w] = a; red[32 + w] = b; }
  __syncthreads();
  float x = l < nw ? red[l] : (MAX0 ? -3.0e38f : 0.f);
  float y = l < nw ? red[32 + l] : 0.f;
  a = MAX0 ? warp_max(x) : warp_sum(x);
  b = warp_sum(y);
}

union V8 {
  uint4 u;
  __half2 h2[4];
  __half h[8];
};

OB_DEVICE uint2 pack8_i8(const float (&f)[8], float s) {
  uint32_t b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (uint32_t)(uint8_t)f2i8_rni_sat(f[i] * s);
  uint2 r;
  r.x = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
  r.y = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
  return r;
}

// ------------------------------------------------------------------------------------------------
// Tensor-parallel all-reduce fused into the norm that consumes it (no reference counterpart: the reference has no
// TP).  The row-parallel GEMM (o_proj / down_proj) of every rank leaves its partial sums in a symmetric buffer that
// all GPUs of the node map over NVLink; the norm kernel of every rank then (1) tells its peers, block by block, that
// its partials are complete and waits for theirs (monotonic epoch flags in peer memory, no reset, CUDA-graph safe),
// (2) reads the W partial rows straight from peer memory, sums them in fp32 in rank order (identical result on every
// rank) and (3) carries on as add + norm (+ quant) -- one kernel instead of NCCL all-reduce + add + norm + quant.
// ------------------------------------------------------------------------------------------------
struct PeerCtx {
  const __half* bufs[8];   // rank p's partial sums [T, H], as mapped in this process
  uint32_t* flags[8];      // rank p's flag array [blocks][8], as mapped in this process
  uint32_t* epoch;         // local: calls made so far, per block
  int world, rank;
};

struct NoPeer {};   // kernel-parameter placeholder of the non-TP instantiations (keeps their parameter block small)
template <bool PEER> struct PeerSel { using type = NoPeer; };
template <> struct PeerSel<true> { using type = PeerCtx; };

// Flag traffic of the exchange barrier.  The partial sums a flag announces were written by the PREVIOUS kernel (the
// row-parallel GEMM), whose completion -- awaited with griddepcontrol.wait before the flag is sent -- already made them
// visible in this GPU's L2, which is where peers read them over NVLink; so the flag itself is a relaxed system-scope store
// and the wait a relaxed system-scope poll followed by ONE acquire fence (a release / acquire pair per poll iteration
// cost ~6 us per exchange at tp2, profiles/r2_tp_step.md).
OB_DEVICE void st_flag_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
OB_DEVICE uint32_t ld_flag_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
OB_DEVICE uint4 ld_peer_v4(const void* p) {   // peer memory is not coherent with this SM's L1: bypass it
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// Block b of every rank: "my partial sums are complete" -> all peers; wait for the same from all peers.  Must run after
// griddepcontrol.wait (the producing GEMM of this rank has finished and flushed).
OB_DEVICE void peer_barrier(const PeerCtx& c, uint32_t* e_s) {
  if (threadIdx.x == 0) {
    const uint32_t e = c.epoch[blockIdx.x] + 1;
    c.epoch[blockIdx.x] = e;
    *e_s = e;
  }
  __syncthreads();
  const uint32_t e = *e_s;
  if ((int)threadIdx.x < c.world) {
    const int p = threadIdx.x;
    st_flag_sys_u32(c.flags[p] + blockIdx.x * 8 + c.rank, e);
    while ((int)(ld_flag_sys_u32(c.flags[c.rank] + blockIdx.x * 8 + p) - e) < 0) {
    }
    asm volatile("fence.acquire.sys;" ::: "memory");
  }
  __syncthreads();
}
// fp32 sum over ranks 0..W-1 of the 8 halves at vector `idx` of `row`, rounded once to fp16
OB_DEVICE uint4 peer_sum_v8(const PeerCtx& c, size_t row, int H, int idx) {
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int p = 0; p < c.world; ++p) {
    V8 x;
    x.u = ld_peer_v4(reinterpret_cast<const uint4*>(c.bufs[p] + row * H) + idx);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __half2float(x.h[j]);
  }
  V8 r;
#pragma unroll
  for (int j = 0; j < 8; ++j) r.h[j] = __float2half_rn(acc[j]);
  return r.u;
}

// All of this thread's vectors of the row at once: the loads of up to four peers x NV vectors are issued back to back
// before anything depends on them, so a row costs ceil(W / 4) NVLink round trips instead of NV x W dependent ones
// (round 2: the one-vector-at-a-time loop made the fused exchange 18 us at tp2, profiles/r2_tp_step.md).  Same fp32
// summation order (rank 0..W-1) and single fp16 rounding as peer_sum_v8.
template <int NV>
OB_DEVICE void peer_sum_row(const PeerCtx& c, size_t row, int H, int nvec, V8 (&out)[NV]) {
  float acc[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int p0 = 0; p0 < c.world; p0 += 4) {
    V8 x[4][NV];
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) {
      if (p0 + pp < c.world) {
        const uint4* base = reinterpret_cast<const uint4*>(c.bufs[p0 + pp] + row * H);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int idx = threadIdx.x + i * blockDim.x;
          if (idx < nvec) x[pp][i].u = ld_peer_v4(base + idx);
        }
      }
    }
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) {
      if (p0 + pp < c.world) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int idx = threadIdx.x + i * blockDim.x;
          if (idx < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] += __half2float(x[pp][i].h[j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) out[i].h[j] = __float2half_rn(acc[i][j]);
}

// ------------------------------------------------------------------------------------------------
// invoke_quant / invoke_quant_fuse_sum   (fused_kernels.cu:57-142)
// ------------------------------------------------------------------------------------------------
template <bool FUSE_SUM>
__global__ void __launch_bounds__(512) quant_kernel(const __half* __restrict__ in, int8_t* __restrict__ out, __half* __restrict__ scale,
                             __half* __restrict__ sum, int H) {
  __shared__ float red[64];
  pdl_trigger();
  pdl_wait();
  const size_t row = blockIdx.x;
  const int nvec = H >> 3;
  const uint4* src = reinterpret_cast<const uint4*>(in + row * H);
  V8 v[MAXV];
  float amax = 0.f, s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = threadIdx.x + i * blockDim.x;
    if (idx < nvec) {
      v[i].u = ld_nc_v4(src + idx);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float f = __half2float(v[i].h[j]);
        s += f;
        amax = fmaxf(amax, fabsf(f));
      }
    }
  }
  block_reduce2<true>(amax, s, red);
  if (threadIdx.x == 0) {
    scale[row] = __float2half_rn(__fdividef(amax, 127.0f));
    if (FUSE_SUM) sum[row] = __float2half_rn(s);
  }
  const float qs = __fdividef(127.0f, amax);
  uint2* dst = reinterpret_cast<uint2*>(out + row * H);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = threadIdx.x + i * blockDim.x;
    if (idx < nvec) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __half2float(v[i].h[j]);
      dst[idx] = pack8_i8(f, qs);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// rms_norm_general(_fuse_sum)  (layernorm_kernels.cu:194-331): (x-mean)*rsqrt(mean(x^2)+eps)*gamma,
// fp16-rounded before amax / sum, per-"reference thread" fp16 partial sums.  blockDim = refblock/8.
// ------------------------------------------------------------------------------------------------
// ADD: x = in + delta (fp16 add, like torch's residual add) is formed first and written to hidden_out.
template <bool FUSE_SUM, bool ADD, bool PEER = false>
__global__ void __launch_bounds__(128) rmsnorm_quant_kernel(const __half* __restrict__ in, const __half* __restrict__ delta,
                                     __half* __restrict__ hidden_out, const __half* __restrict__ gamma,
                                     int8_t* __restrict__ out, __half* __restrict__ scale, __half* __restrict__ sum,
                                     int H, float eps, const typename PeerSel<PEER>::type pc) {
  __shared__ float red[64];
  __shared__ uint32_t epoch_s;
  pdl_trigger();
  const size_t row = blockIdx.x;
  const int nvec = H >> 3;
  const uint4* src = reinterpret_cast<const uint4*>(in + row * H);
  const uint4* gsrc = reinterpret_cast<const uint4*>(gamma);
  V8 v[MAXV];
  float s1 = 0.f, s2 = 0.f;
  pdl_wait();
  [[maybe_unused]] V8 psum[PEER ? 4 : 1];
  if constexpr (PEER) {
    peer_barrier(pc, &epoch_s);
    peer_sum_row<4>(pc, row, H, nvec, psum);   // rows of <= 4096 halves (4 vectors per thread at 128 threads)
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = threadIdx.x + i * blockDim.x;
    if (idx < nvec) {
      v[i].u = ld_nc_v4(src + idx);
      if (ADD) {
        V8 dl;
        if constexpr (PEER) dl.u = (i < 4) ? psum[i < 4 ? i : 0].u : peer_sum_v8(pc, row, H, idx);
        else dl.u = ld_nc_v4(reinterpret_cast<const uint4*>(delta + row * H) + idx);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i].h2[j] = __hadd2(v[i].h2[j], dl.h2[j]);
        reinterpret_cast<uint4*>(hidden_out + row * H)[idx] = v[i].u;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float f = __half2float(v[i].h[j]);
        s1 += f;
        s2 += f * f;
      }
    }
  }
  block_reduce2<false>(s1, s2, red);
  const float mean = __fdividef(s1, (float)H);
  const float rstd = rsqrtf(__fdividef(s2, (float)H) + eps);
  float amax = 1.013279e-06f;  // (half)1e-6f, layernorm_kernels.cu:279
  __half hsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) hsum[j] = __float2half_rn(0.f);
  float nf[MAXV][8];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = threadIdx.x + i * blockDim.x;
    if (idx < nvec) {
      V8 g;
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
  block_reduce2<true>(amax, ps, red);
  if (threadIdx.x == 0) {
    scale[row] = __float2half_rn(__fdividef(amax, 127.0f));
    if (FUSE_SUM) sum[row] = __float2half_rn(ps);
  }
  const float qs = __fdividef(127.0f, amax);
  uint2* dst = reinterpret_cast<uint2*>(out + row * H);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = threadIdx.x + i * blockDim.x;
    if (idx < nvec) dst[idx] = pack8_i8(nf[i], qs);
  }
}

// plain rms_norm, fp16 out (layernorm_kernels.cu:335-364): ((half)(x*rstd)) * w in half
template <bool ADD, bool PEER = false>
__global__ void __launch_bounds__(512) rmsnorm_f16_kernel(const __half* __restrict__ in, const __half* __restrict__ delta,
                                   const __half* __restrict__ gamma, __half* __restrict__ out, int H, float eps,
                                   const typename PeerSel<PEER>::type pc) {
  __shared__ float red[64];
  __shared__ uint32_t epoch_s;
  pdl_trigger();
  const size_t row = blockIdx.x;
  const int nvec = H >> 3;
  const uint4* src = reinterpret_cast<const uint4*>(in + row * H);
  const uint4* gsrc = reinterpret_cast<const uint4*>(gamma);
  V8 v[MAXV];
  float s2 = 0.f, dummy = 0.f;
  pdl_wait();
  if constexpr (PEER) peer_barrier(pc, &epoch_s);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = threadIdx.x + i * blockDim.x;
    if (idx < nvec) {
      v[i].u = ld_nc_v4(src + idx);
      if (ADD) {
        V8 dl;
        if constexpr (PEER) dl.u = peer_sum_v8(pc, row, H, idx);
        else dl.u = ld_nc_v4(reinterpret_cast<const uint4*>(delta + row * H) + idx);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i].h2[j] = __hadd2(v[i].h2[j], dl.h2[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float f = __half2float(v[i].h[j]);
        s2 += f * f;
      }
    }
  }
  block_reduce2<false>(s2, dummy, red);
  const float rstd = rsqrtf(__fdividef(s2, (float)H) + eps);
  uint4* dst = reinterpret_cast<uint4*>(out + row * H);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = threadIdx.x + i * blockDim.x;
    if (idx < nvec) {
      V8 g, o;
      g.u = __ldg(gsrc + idx);
#pragma unroll
      for (int j = 0; j < 8; ++j) o.h[j] = __hmul(__float2half_rn(__half2float(v[i].h[j]) * rstd), g.h[j]);
      dst[idx] = o.u;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// silu_and_mul (activation_kernels.cu:10-30) and the fused silu*mul -> per-token quant(+sum) that
// removes the [T, inter] fp16 round trip of activation.py:54-64.
// ------------------------------------------------------------------------------------------------
OB_DEVICE __half silu_mul(__half g, __half u) {
  const float x = __half2float(g);
  const __half s = __float2half_rn(__fdividef(x, 1.0f + __expf(-x)));
  return __hmul(s, u);
}

__global__ void __launch_bounds__(512) silu_and_mul_kernel(const __half* __restrict__ in, __half* __restrict__ out, int d) {
  pdl_trigger();
  pdl_wait();
  const size_t row = blockIdx.x;
  const uint4* g = reinterpret_cast<const uint4*>(in + row * 2 * d);
  const uint4* u = reinterpret_cast<const uint4*>(in + row * 2 * d + d);
  uint4* dst = reinterpret_cast<uint4*>(out + row * d);
  for (int idx = threadIdx.x; idx < (d >> 3); idx += blockDim.x) {
    V8 a, b, o;
    a.u = ld_nc_v4(g + idx);
    b.u = ld_nc_v4(u + idx);
#pragma unroll
    for (int j = 0; j < 8; ++j) o.h[j] = silu_mul(a.h[j], b.h[j]);
    dst[idx] = o.u;
  }
}

// NV = 16-byte vectors of the output row cached per thread (compile-time so that the register footprint -- and with
// it the number of rows in flight per SM at prefill sizes -- follows the row length instead of the worst case).
// OCC3: cap registers for three resident CTAs per SM -- pays at prefill sizes (2.4 -> 4.5 TB/s at T = 8192), costs a few
// spilled registers, so decode-sized launches (latency bound, 64 rows) use the uncapped instantiation.
template <bool FUSE_SUM, int NV, bool OCC3>
__global__ void __launch_bounds__(512, OCC3 ? 3 : 1) silu_mul_quant_kernel(const __half* __restrict__ in, int8_t* __restrict__ out,
                                      __half* __restrict__ scale, __half* __restrict__ sum, int d) {
  __shared__ float red[64];
  pdl_trigger();
  pdl_wait();
  const size_t row = blockIdx.x;
  const int nvec = d >> 3;
  const uint4* g = reinterpret_cast<const uint4*>(in + row * 2 * d);
  const uint4* u = reinterpret_cast<const uint4*>(in + row * 2 * d + d);
  V8 v[NV];
  float amax = 0.f, s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = threadIdx.x + i * blockDim.x;
    if (idx < nvec) {
      V8 a, b;
      a.u = ld_nc_v4(g + idx);
      b.u = ld_nc_v4(u + idx);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i].h[j] = silu_mul(a.h[j], b.h[j]);
        const float f = __half2float(v[i].h[j]);
        s += f;
        amax = fmaxf(amax, fabsf(f));
      }
    }
  }
  block_reduce2<true>(amax, s, red);
  if (threadIdx.x == 0) {
    scale[row] = __float2half_rn(__fdividef(amax, 127.0f));
    if (FUSE_SUM) sum[row] = __float2half_rn(s);
  }
  const float qs = __fdividef(127.0f, amax);
  uint2* dst = reinterpret_cast<uint2*>(out + row * d);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = threadIdx.x + i * blockDim.x;
    if (idx < nvec) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __half2float(v[i].h[j]);
      dst[idx] = pack8_i8(f, qs);
    }
  }
}

// fp16 residual add (llama_w4a8_unpad.py:425,437 `residual + out_down_proj_act_buffer`)
__global__ void add_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ out,
                           size_t nvec) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    V8 x, y, o;
    x.u = reinterpret_cast<const uint4*>(a)[i];
    y.u = reinterpret_cast<const uint4*>(b)[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) o.h2[j] = __hadd2(x.h2[j], y.h2[j]);
    reinterpret_cast<uint4*>(out)[i] = o.u;
  }
}

// ---------------------------------------------------------------------------------------------- host
// <= 512 threads per row (kernels are compiled with __launch_bounds__(512)); up to MAXV vectors per thread
static int pick_threads(int nvec, int maxv) {
  int t = (nvec + maxv - 1) / maxv;
  t = ((t + 31) / 32) * 32;
  if (t > 512) {
    t = (nvec + MAXV - 1) / MAXV;
    t = ((t + 31) / 32) * 32;
  }
  if (t < 32) t = 32;
  return t;
}
static int check(int H, int maxthreads = 512) {
  if (H <= 0 || (H & 7)) return OB_ERR_SHAPE;
  if ((H >> 3) > MAXV * maxthreads) return OB_ERR_SHAPE;
  return 0;
}
#define OB_LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? 0 : OB_ERR_CUDA)

int quant_run(const __half* in, int8_t* out, __half* scale, __half* sum, int T, int H, cudaStream_t st) {
  if (T <= 0) return 0;
  if (int e = check(H)) return e;
  const int th = pick_threads(H >> 3, 4);
  cudaError_t e = sum ? launch_pdl(quant_kernel<true>, dim3(T), dim3(th), 0, st, in, out, scale, sum, H)
                      : launch_pdl(quant_kernel<false>, dim3(T), dim3(th), 0, st, in, out, scale, (__half*)nullptr, H);
  return e == cudaSuccess ? 0 : OB_ERR_CUDA;
}

int rmsnorm_quant_run(const __half* in, const __half* delta, __half* hidden_out, const __half* gamma, int8_t* out,
                      __half* scale, __half* sum, int T, int H, float eps, cudaStream_t st) {
  if (T <= 0) return 0;
  // reference block = min(H,1024) rounded up to 32 threads, one element per thread per iteration
  int refblock = std::min(H, 1024);
  refblock = 32 * ((refblock + 31) / 32);
  if (int e = check(H, refblock / 8)) return e;
  if (H % refblock != 0 && H > refblock) return OB_ERR_SHAPE;
  const int th = std::max(32, refblock / 8);
  cudaError_t e;
  const dim3 g(T), b(th);
  if (delta) {
    e = sum ? launch_pdl(rmsnorm_quant_kernel<true, true>, g, b, 0, st, in, delta, hidden_out, gamma, out, scale, sum, H, eps, NoPeer{})
            : launch_pdl(rmsnorm_quant_kernel<false, true>, g, b, 0, st, in, delta, hidden_out, gamma, out, scale, sum, H, eps, NoPeer{});
  } else {
    e = sum ? launch_pdl(rmsnorm_quant_kernel<true, false>, g, b, 0, st, in, delta, hidden_out, gamma, out, scale, sum, H, eps, NoPeer{})
            : launch_pdl(rmsnorm_quant_kernel<false, false>, g, b, 0, st, in, delta, hidden_out, gamma, out, scale, sum, H, eps, NoPeer{});
  }
  return e == cudaSuccess ? 0 : OB_ERR_CUDA;
}

int rmsnorm_f16_run(const __half* in, const __half* delta, const __half* gamma, __half* out, int T, int H, float eps,
                    cudaStream_t st) {
  if (T <= 0) return 0;
  if (int e = check(H)) return e;
  const dim3 g(T), b(pick_threads(H >> 3, 4));
  cudaError_t e = delta ? launch_pdl(rmsnorm_f16_kernel<true>, g, b, 0, st, in, delta, gamma, out, H, eps, NoPeer{})
                        : launch_pdl(rmsnorm_f16_kernel<false>, g, b, 0, st, in, delta, gamma, out, H, eps, NoPeer{});
  return e == cudaSuccess ? 0 : OB_ERR_CUDA;
}

static int fill_peer(PeerCtx& pc, const PeerArgs& a, int T) {
  if (a.world < 2 || a.world > 8 || a.rank < 0 || a.rank >= a.world || !a.epoch || T > a.max_blocks) return OB_ERR_ARG;
  for (int p = 0; p < a.world; ++p) {
    if (!a.bufs[p] || !a.flags[p]) return OB_ERR_ARG;
    pc.bufs[p] = a.bufs[p];
    pc.flags[p] = a.flags[p];
  }
  pc.epoch = a.epoch; pc.world = a.world; pc.rank = a.rank;
  return 0;
}

int peer_rmsnorm_quant_run(const __half* in, const PeerArgs& peer, __half* hidden_out, const __half* gamma, int8_t* out,
                           __half* scale, __half* sum, int T, int H, float eps, cudaStream_t st) {
  if (T <= 0) return 0;
  int refblock = std::min(H, 1024);
  refblock = 32 * ((refblock + 31) / 32);
  if (int e = check(H, refblock / 8)) return e;
  if (H % refblock != 0 && H > refblock) return OB_ERR_SHAPE;
  PeerCtx pc{};
  if (int e = fill_peer(pc, peer, T)) return e;
  const dim3 g(T), b(std::max(32, refblock / 8));
  const __half* nodelta = nullptr;
  cudaError_t e = sum ? launch_pdl(rmsnorm_quant_kernel<true, true, true>, g, b, 0, st, in, nodelta, hidden_out, gamma, out, scale, sum, H, eps, pc)
                      : launch_pdl(rmsnorm_quant_kernel<false, true, true>, g, b, 0, st, in, nodelta, hidden_out, gamma, out, scale, sum, H, eps, pc);
  return e == cudaSuccess ? 0 : OB_ERR_CUDA;
}

int peer_rmsnorm_f16_run(const __half* in, const PeerArgs& peer, const __half* gamma, __half* out, int T, int H, float eps,
                         cudaStream_t st) {
  if (T <= 0) return 0;
  if (int e = check(H)) return e;
  PeerCtx pc{};
  if (int e = fill_peer(pc, peer, T)) return e;
  const dim3 g(T), b(pick_threads(H >> 3, 4));
  const __half* nodelta = nullptr;
  return launch_pdl(rmsnorm_f16_kernel<true, true>, g, b, 0, st, in, nodelta, gamma, out, H, eps, pc) == cudaSuccess ? 0 : OB_ERR_CUDA;
}

int silu_and_mul_run(const __half* in, __half* out, int T, int d, cudaStream_t st) {
  if (T <= 0) return 0;
  if (d <= 0 || (d & 7)) return OB_ERR_SHAPE;
  return launch_pdl(silu_and_mul_kernel, dim3(T), dim3(pick_threads(d >> 3, 4)), 0, st, in, out, d) == cudaSuccess
             ? 0 : OB_ERR_CUDA;
}

int silu_mul_quant_run(const __half* in, int8_t* out, __half* scale, __half* sum, int T, int d, cudaStream_t st) {
  if (T <= 0) return 0;
  if (int e = check(d)) return e;
  const int th = pick_threads(d >> 3, 4);
  const int nv = ((d >> 3) + th - 1) / th;
  cudaError_t e;
#define OB_SILU(NVV, OCC)                                                                                              \
  e = sum ? launch_pdl(silu_mul_quant_kernel<true, NVV, OCC>, dim3(T), dim3(th), 0, st, in, out, scale, sum, d)        \
          : launch_pdl(silu_mul_quant_kernel<false, NVV, OCC>, dim3(T), dim3(th), 0, st, in, out, scale, (__half*)nullptr, d)
  const bool big = T >= 1024;
  if (nv <= 1) { OB_SILU(1, false); } else if (nv <= 2) { OB_SILU(2, false); }
  else if (nv <= 4) { if (big) { OB_SILU(4, true); } else { OB_SILU(4, false); } }
  else { OB_SILU(8, false); }
#undef OB_SILU
  return e == cudaSuccess ? 0 : OB_ERR_CUDA;
}

int add_run(const __half* a, const __half* b, __half* out, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  if (n & 7) return OB_ERR_SHAPE;
  const size_t nvec = n >> 3;
  const int blocks = (int)std::min<size_t>((nvec + 255) / 256, 148 * 8);
  return launch_pdl(add_kernel, dim3(blocks), dim3(256), 0, st, a, b, out, nvec) == cudaSuccess ? 0 : OB_ERR_CUDA;
}

}  // namespace ob
