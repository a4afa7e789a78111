// LServe page statistics and page selector for sm_100a.
//
// Replaces (same results, new design):
//   paged_min_max_pool          /root/reference/kernels/csrc/fused_attention/sparse_utils/ContextPool/
//                               context_pool_kernel.cu:16-69 (kernel), :145-213 (host op)
//   single_query_page_selector  .../sparse_utils/KVPageSelector/KVPageSelectorTemplate.hpp:478-503 (score),
//                               :786-1290 (kernel), fused_kv_page_selector.cpp:171-334 (host op)
//
// Page layout (common/kvCacheUtils.h:53-164, cache_engine.py:73-88), K page of the retrieval pool:
//   [H_pool][64 tok][Dh/2] nibbles | scales f16 [H_pool][64] | zeros f16 [H_pool][64] |
//   kmax f16 [64/sub][H_pool*Dh] | kmin f16 [64/sub][H_pool*Dh]
//
// Both ops are pure HBM streams (min/max and fp16 multiply-add on CUDA cores):
//   * pool: one warp per (sequence, pooled head, 16-token sub-chunk) reads the 16 x 256 B key rows with
//     16-byte loads that are all in flight at once (the reference reduces with 4 x 8 x log2(16) half shuffles
//     per 16 B and one token per thread);
//   * selector: one half-warp per page; the kmax / kmin rows of its four sub-chunks (8 x 256 B) are read ONCE per
//     kv head and scored against all query heads of the GQA group (the reference launches one CTA per query
//     head and re-reads the statistics for each).  The fp16 arithmetic of the score follows the reference
//     operation for operation, including the 16-lane fp32 butterfly, so scores are bit-identical given the
//     same rotated q.
#include "kv4_attention.h"
#include "launch.h"
#include "lserve_ops.h"
#include "ptx.cuh"

#include <algorithm>

namespace ob {

constexpr int LS_DH = 128;

// ------------------------------------------------------------------------------------------------ pool
struct PoolParams {
  const __half* keys;            // [T, H_in, Dh] post-RoPE keys
  long long row_stride, head_stride;  // elements
  const int64_t* tab;            // [B, 2, max_pages]
  int max_pages;
  const int* cu_seqlens;         // [B+1]
  const int* pooling_heads_idx;  // [pool_h] -> input head
  int pool_h, sub, page_size;
  long long stats_off_bytes;     // from the K page base to kmax
  int eles_per_indicator;        // pool_h * Dh
};

__global__ void __launch_bounds__(256) paged_min_max_pool_kernel(const PoolParams p) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, ph = blockIdx.z;
  const int s0 = p.cu_seqlens[b];
  const int L = p.cu_seqlens[b + 1] - s0;
  const int chunk = blockIdx.x * 8 + warp;      // sub-chunk index inside the sequence
  const int t0 = chunk * p.sub;
  if (t0 >= L) return;
  const int n_tok = min(p.sub, L - t0);
  const int in_head = p.pooling_heads_idx[ph];
  // half-warp hw takes tokens hw, hw+2, ...; lane (l & 15) owns channels 8*(l&15) .. +7 (one uint4)
  const int hw = lane >> 4, cl = lane & 15;
  const __half* base = p.keys + (size_t)(s0 + t0) * p.row_stride + (size_t)in_head * p.head_stride + cl * 8;
  uint4 v[16];
  const int n_mine = (n_tok - hw + 1) >> 1;      // tokens hw, hw+2, ... < n_tok
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < n_mine) v[i] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(hw + 2 * i) * p.row_stride));
  __half2 mx[4], mn[4];
  bool have = n_mine > 0;
  if (have) {
    const __half2* h = reinterpret_cast<const __half2*>(&v[0]);
#pragma unroll
    for (int j = 0; j < 4; ++j) { mx[j] = h[j]; mn[j] = h[j]; }
  }
#pragma unroll
  for (int i = 1; i < 16; ++i) {
    if (i < n_mine) {
      const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) { mx[j] = __hmax2(mx[j], h[j]); mn[j] = __hmin2(mn[j], h[j]); }
    }
  }
  // combine the two half-warps (the odd half may own no token when n_tok == 1)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t a = *reinterpret_cast<uint32_t*>(&mx[j]), c = *reinterpret_cast<uint32_t*>(&mn[j]);
    const uint32_t oa = __shfl_xor_sync(0xffffffffu, a, 16), oc = __shfl_xor_sync(0xffffffffu, c, 16);
    const bool other_has = __shfl_xor_sync(0xffffffffu, (int)have, 16) != 0;
    if (have && other_has) {
      mx[j] = __hmax2(mx[j], *reinterpret_cast<const __half2*>(&oa));
      mn[j] = __hmin2(mn[j], *reinterpret_cast<const __half2*>(&oc));
    } else if (!have && other_has) {
      mx[j] = *reinterpret_cast<const __half2*>(&oa);
      mn[j] = *reinterpret_cast<const __half2*>(&oc);
    }
  }
  if (hw == 0) {
    const int page = t0 / p.page_size;
    const int sub_idx = (t0 % p.page_size) / p.sub;
    uint8_t* kpage = reinterpret_cast<uint8_t*>(p.tab[(size_t)b * 2 * p.max_pages + page]);
    __half* kmax = reinterpret_cast<__half*>(kpage + p.stats_off_bytes) + (size_t)sub_idx * p.eles_per_indicator +
                   (size_t)ph * p.head_stride + cl * 8;
    __half* kmin = kmax + (size_t)(p.page_size / p.sub) * p.eles_per_indicator;
    *reinterpret_cast<uint4*>(kmax) = *reinterpret_cast<uint4*>(mx);
    *reinterpret_cast<uint4*>(kmin) = *reinterpret_cast<uint4*>(mn);
  }
}

int paged_min_max_pool_run(const PoolArgs& a, cudaStream_t st) {
  if (a.batch <= 0 || a.num_pooling_heads <= 0 || a.max_seqlen <= 0) return 0;
  if (a.head_dim != LS_DH) return OB_ERR_SHAPE;
  if (a.pooling_size <= 0 || a.pooling_size > 32 || a.page_size % a.pooling_size) return OB_ERR_SHAPE;
  if (a.head_stride % 8 || a.row_stride % 8) return OB_ERR_ALIGN;
  PoolParams p{};
  p.keys = a.keys; p.row_stride = a.row_stride; p.head_stride = a.head_stride;
  p.tab = a.retrieval_kv_pointers; p.max_pages = a.r_max_pages;
  p.cu_seqlens = a.cu_seqlens; p.pooling_heads_idx = a.pooling_heads_idx;
  p.pool_h = a.num_pooling_heads; p.sub = a.pooling_size; p.page_size = a.page_size;
  // context_pool_kernel.cu:45: (k_cache + mBytesPerSeq) as half* + tokens_per_block * pool_h * (zeros ? 2 : 1)
  p.stats_off_bytes = (long long)a.page_size * a.size_per_retrieval_token +
                      (long long)a.page_size * a.num_pooling_heads * (a.kv_cache_with_zeros ? 2 : 1) * 2;
  p.eles_per_indicator = a.num_pooling_heads * a.head_dim;
  const int chunks = (a.max_seqlen + a.pooling_size - 1) / a.pooling_size;
  dim3 grid((chunks + 7) / 8, a.batch, a.num_pooling_heads);
  return launch_pdl(paged_min_max_pool_kernel, grid, dim3(256), 0, st, p) == cudaSuccess ? 0 : OB_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------ selector
struct SelParams {
  const __half* q; long long q_bs;       // [B, Hq, Dh] view, head stride Dh
  __half* out;                           // [B, Hq, padded_host] zero-initialised by the caller
  long long out_elems;
  const int64_t* tab; int max_pages;     // retrieval [B,2,max_pages]
  const int* lengths; int timestep;
  const int* retrieval_flags; const int* head_rank;
  int B, Hq, Hkv, hpool;
  int sub, group_size;                   // tokens per sub-chunk, sub-chunks per page
  long long stats_off_bytes; int eles_per_indicator;
  float rope_base, rope_scale; int rotary_dim;
};

// Score of one sub-chunk for one query head, held by one half-warp: lane (l & 15) owns channels 8*(l&15)..+7.
// Mirrors qk_hmma_dot_min_max<16> (KVPageSelectorTemplate.hpp:478-503) operation for operation.
OB_DEVICE float subchunk_score(const uint4& q, const uint4& kmax, const uint4& kmin) {
  const __half2* qp = reinterpret_cast<const __half2*>(&q);
  const __half2* xp = reinterpret_cast<const __half2*>(&kmax);
  const __half2* np_ = reinterpret_cast<const __half2*>(&kmin);
  __half2 acc = __hmax2(__hmul2(qp[0], xp[0]), __hmul2(qp[0], np_[0]));
#pragma unroll
  for (int i = 1; i < 4; ++i) acc = __hadd2(acc, __hmax2(__hmul2(qp[i], xp[i]), __hmul2(qp[i], np_[i])));
  float s = __half2float(__hadd(__low2half(acc), __high2half(acc)));
#pragma unroll
  for (int m = 8; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  return s;
}

template <int G>
__global__ void __launch_bounds__(256) page_selector_kernel(const SelParams p) {
  __shared__ __align__(16) __half q_s[G][LS_DH];
  __shared__ float rope_cs[LS_DH / 2], rope_sn[LS_DH / 2];
  pdl_trigger();
  pdl_wait();   // lengths / q (and, in principle, the tables) may be the previous kernel's output: read nothing before
  const int b = blockIdx.z, hkv = blockIdx.y;
  if (p.retrieval_flags && p.retrieval_flags[hkv] == 0) return;   // streaming heads keep their zero rows
  const int rank = p.head_rank ? p.head_rank[hkv] : hkv;
  const int tid = threadIdx.x;
  const int tl = p.lengths ? p.lengths[b] - 1 : p.timestep;       // cached tokens == position of the query
  const int n_sub = (tl + p.sub - 1) / p.sub;
  const int padded = (n_sub + p.group_size - 1) / p.group_size * p.group_size;   // row pitch used by the reference kernel
  const int n_pages = padded / p.group_size;
  if ((int)blockIdx.x * 16 >= n_pages) return;
  const int half_rot = p.rotary_dim >> 1;
  if (tid < LS_DH / 2 && tid < half_rot) {
    const float inv_freq = ((float)tl * p.rope_scale) / powf(p.rope_base, (float)(2 * tid) / (float)p.rotary_dim);
    sincosf(inv_freq, &rope_sn[tid], &rope_cs[tid]);
  }
  __syncthreads();
  const int hq0 = hkv * G;
  for (int item = tid; item < G * (LS_DH / 2); item += 256) {
    const int h = item / (LS_DH / 2), d = item - h * (LS_DH / 2);
    const __half* src = p.q + (size_t)b * p.q_bs + (size_t)(hq0 + h) * LS_DH;
    if (d < half_rot) {
      const float sn = rope_sn[d], cs = rope_cs[d];
      const float x = __half2float(src[d]), y = __half2float(src[d + half_rot]);
      q_s[h][d] = __float2half_rn(cs * x - sn * y);
      q_s[h][d + half_rot] = __float2half_rn(cs * y + sn * x);
    } else {
      const int e = p.rotary_dim + 2 * (d - half_rot);
      q_s[h][e] = src[e];
      q_s[h][e + 1] = src[e + 1];
    }
  }
  __syncthreads();
  const int hw = tid >> 4, cl = tid & 15;
  const int page = blockIdx.x * 16 + hw;
  // every lane of a warp runs the shuffles; lanes of a page beyond the end compute on page n_pages-1 and do not store
  const int page_c = min(page, n_pages - 1);
  const uint8_t* kpage = reinterpret_cast<const uint8_t*>(p.tab[(size_t)b * 2 * p.max_pages + page_c]);
  const __half* kmax0 = reinterpret_cast<const __half*>(kpage + p.stats_off_bytes) + (size_t)rank * LS_DH + cl * 8;
  const __half* kmin0 = kmax0 + (size_t)p.group_size * p.eles_per_indicator;
  uint4 qv[G];
#pragma unroll
  for (int h = 0; h < G; ++h) qv[h] = *reinterpret_cast<const uint4*>(&q_s[h][cl * 8]);
  for (int s0 = 0; s0 < p.group_size; s0 += 4) {
    uint4 kx[4], kn[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s0 + s < p.group_size) {
        kx[s] = __ldg(reinterpret_cast<const uint4*>(kmax0 + (size_t)(s0 + s) * p.eles_per_indicator));
        kn[s] = __ldg(reinterpret_cast<const uint4*>(kmin0 + (size_t)(s0 + s) * p.eles_per_indicator));
      }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s0 + s >= p.group_size) break;
      const int sc = page_c * p.group_size + s0 + s;
#pragma unroll
      for (int h = 0; h < G; ++h) {
        const float v = subchunk_score(qv[h], kx[s], kn[s]);
        if (cl == 0 && page < n_pages && sc < n_sub) {
          const long long idx = ((long long)b * p.Hq + hq0 + h) * padded + sc;
          if (idx < p.out_elems) p.out[idx] = __float2half(v);
        }
      }
    }
  }
}

int page_selector_run(const SelectorArgs& a, cudaStream_t st) {
  if (a.B <= 0) return 0;
  if (a.head_dim != LS_DH || a.Hq % a.Hkv) return OB_ERR_SHAPE;
  if (a.tokens_per_sub_chunk <= 0 || a.tokens_per_block % a.tokens_per_sub_chunk) return OB_ERR_SHAPE;
  const int G = a.Hq / a.Hkv;
  SelParams p{};
  p.q = a.q; p.q_bs = a.q_bs; p.out = a.out;
  p.tab = a.retrieval_kv_pointers; p.max_pages = a.r_max_pages;
  p.lengths = a.lengths; p.timestep = a.timestep;
  p.retrieval_flags = a.retrieval_head_flags; p.head_rank = a.head_rank_table;
  p.B = a.B; p.Hq = a.Hq; p.Hkv = a.Hkv; p.hpool = a.num_retrieval_kv_heads;
  p.sub = a.tokens_per_sub_chunk; p.group_size = a.tokens_per_block / a.tokens_per_sub_chunk;
  p.stats_off_bytes = (long long)a.tokens_per_block * a.size_per_retrieval_token +
                      (long long)a.tokens_per_block * a.num_retrieval_kv_heads * 2 * 2;
  p.eles_per_indicator = a.hidden_dim_per_retrieval_token;
  p.rope_base = a.rotary_base; p.rope_scale = a.rotary_scale; p.rotary_dim = a.rotary_dim;
  // fused_kv_page_selector.cpp:274-277: the host sizes the output from `timestep`
  const int n_sub_host = (a.timestep + p.sub - 1) / p.sub;
  const int padded_host = (n_sub_host + p.group_size - 1) / p.group_size * p.group_size;
  p.out_elems = (long long)a.B * a.Hq * padded_host;
  if (p.out_elems == 0) return 0;
  if (cudaMemsetAsync(a.out, 0, (size_t)p.out_elems * 2, st) != cudaSuccess) return OB_ERR_CUDA;
  const int max_pages = padded_host / p.group_size;
  dim3 grid((max_pages + 15) / 16, a.Hkv, a.B);
#define OB_SEL(g) \
  case g: return launch_pdl(page_selector_kernel<g>, grid, dim3(256), 0, st, p) == cudaSuccess ? 0 : OB_ERR_CUDA;
  switch (G) {
    OB_SEL(1) OB_SEL(2) OB_SEL(4) OB_SEL(8)
    default: return OB_ERR_SHAPE;
  }
#undef OB_SEL
}

// ------------------------------------------------------------------------------------------------ page top-k
// The page choice the reference makes in PyTorch after the selector (decoding_attention.py:132-141):
//   page score = max over the page's 4 sub-chunk scores; keep the k-1 best pages among all but the newest; append the
//   newest page; int32.  One CTA per (sequence, q-head): fp16 scores are mapped to order-preserving 16-bit keys and the
//   (k-1)-th largest is found with two 256-bin radix passes over shared memory; pages strictly above the threshold are
//   emitted first, ties at the threshold in page order (torch.topk's tie order is unspecified; the attention only needs the
//   SET and the newest page last).  Output order: ascending page index within each class -- deterministic.
constexpr int TOPK_THREADS = 256;

OB_DEVICE uint32_t f16_key(__half h) {   // order-preserving map fp16 -> [0, 65535]; -0 < +0, NaNs sort to the ends
  const uint32_t u = __half_as_ushort(h);
  return (u & 0x8000u) ? (0xFFFFu - u) : (u | 0x8000u);
}

__global__ void __launch_bounds__(TOPK_THREADS) page_topk_kernel(const __half* __restrict__ scores, int* __restrict__ out,
                                                                   const int pitch /*sub-chunks per row*/, const int group /*per page*/,
                                                                   const int total_pages, const int k_out) {
  extern __shared__ uint16_t keys[];          // [total_pages - 1]
  __shared__ int hist[256];
  __shared__ int sel[4];                      // threshold bin bookkeeping
  __shared__ int warp_cnt[2][TOPK_THREADS / 32];
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __half* src = scores + (size_t)row * pitch;
  int* dst = out + (size_t)row * k_out;
  const int n = total_pages - 1;              // candidates: every page but the newest
  const int k = k_out - 1;                    // how many of them to keep
  if (tid == 0) dst[k_out - 1] = total_pages - 1;
  if (k <= 0) return;
  for (int pg = tid; pg < n; pg += TOPK_THREADS) {
    __half m = src[pg * group];
    for (int j = 1; j < group; ++j) m = __hmax(m, src[pg * group + j]);
    keys[pg] = (uint16_t)f16_key(m);
  }
  // pass 1: high byte
  for (int i = tid; i < 256; i += TOPK_THREADS) hist[i] = 0;
  __syncthreads();
  for (int pg = tid; pg < n; pg += TOPK_THREADS) atomicAdd(&hist[keys[pg] >> 8], 1);
  __syncthreads();
  if (tid == 0) {
    int acc = 0, b = 255;
    for (; b >= 0; --b) { if (acc + hist[b] >= k) break; acc += hist[b]; }
    sel[0] = b; sel[1] = acc;                 // `acc` keys lie in bins above b
  }
  __syncthreads();
  const int hb = sel[0], above_hi = sel[1];
  __syncthreads();
  // pass 2: low byte inside the threshold bin
  for (int i = tid; i < 256; i += TOPK_THREADS) hist[i] = 0;
  __syncthreads();
  for (int pg = tid; pg < n; pg += TOPK_THREADS)
    if ((keys[pg] >> 8) == hb) atomicAdd(&hist[keys[pg] & 0xFF], 1);
  __syncthreads();
  if (tid == 0) {
    int acc = above_hi, b = 255;
    for (; b >= 0; --b) { if (acc + hist[b] >= k) break; acc += hist[b]; }
    sel[2] = (hb << 8) | b;                   // threshold key
    sel[3] = acc;                             // keys strictly above the threshold
  }
  __syncthreads();
  const int thr = sel[2], n_above = sel[3];
  // emit: pages above the threshold first, then ties in page order; slots from a block-wide exclusive scan
  int base_a = 0, base_t = 0;
  for (int p0 = 0; p0 < n; p0 += TOPK_THREADS) {
    const int pg = p0 + tid;
    const int key = pg < n ? keys[pg] : -1;
    const bool is_a = key > thr, is_t = key == thr;
    const unsigned ma = __ballot_sync(0xffffffffu, is_a), mt = __ballot_sync(0xffffffffu, is_t);
    if (lane == 0) { warp_cnt[0][warp] = __popc(ma); warp_cnt[1][warp] = __popc(mt); }
    __syncthreads();
    int off_a = base_a, off_t = base_t, tot_a = 0, tot_t = 0;
    for (int w = 0; w < TOPK_THREADS / 32; ++w) {
      if (w < warp) { off_a += warp_cnt[0][w]; off_t += warp_cnt[1][w]; }
      tot_a += warp_cnt[0][w]; tot_t += warp_cnt[1][w];
    }
    const unsigned lt = (1u << lane) - 1u;
    if (is_a) dst[off_a + __popc(ma & lt)] = pg;
    if (is_t) {
      const int slot = n_above + off_t + __popc(mt & lt);
      if (slot < k) dst[slot] = pg;
    }
    base_a += tot_a; base_t += tot_t;
    __syncthreads();
  }
}

int page_topk_run(const __half* scores, int* out, int rows, int pitch_sub_chunks, int sub_chunks_per_page, int total_pages,
                  int k_out, cudaStream_t st) {
  if (rows <= 0) return 0;
  if (total_pages <= 0 || k_out <= 0 || k_out > total_pages || sub_chunks_per_page <= 0 ||
      pitch_sub_chunks < total_pages * sub_chunks_per_page)
    return OB_ERR_SHAPE;
  const size_t smem = (size_t)std::max(1, total_pages - 1) * 2;
  if (smem > 200 * 1024) return OB_ERR_SHAPE;   // > 100K pages (6.5M tokens)
  static bool attr_done[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 16 && !attr_done[dev] && smem > 48 * 1024) {
    if (cudaFuncSetAttribute(page_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return OB_ERR_CUDA;
    attr_done[dev] = true;
  }
  return launch_pdl(page_topk_kernel, dim3(rows), dim3(TOPK_THREADS), smem, st, scores, out, pitch_sub_chunks, sub_chunks_per_page,
                    total_pages, k_out) == cudaSuccess ? 0 : OB_ERR_CUDA;
}

}  // namespace ob
