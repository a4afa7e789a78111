// KV4 paged-attention decode for sm_100a (QServe dense + LServe streaming / dynamic page-select masks).
//
// Replaces (same semantics, new design):
//   /root/reference/kernels/csrc/fused_attention/fused_attention_pure_dense/
//       decoderMaskedMultiheadAttentionTemplate.hpp:743-2222   (QServe dense)
//   .../fused_attention_fine_grained/{dense,sparse}_attention/decoderMaskedMultiheadAttentionTemplate.hpp
//       (LServe: retrieval/streaming head split, ring pages, dynamic page redirect, multi-block)
//
// Design differences from the reference (which launches one CTA per *query* head and keeps every
// logit of the context in shared memory):
//   * one CTA per (sequence, KV head, KV split) serves the whole GQA group, so K/V bytes are read
//     from HBM once per KV head instead of once per query head;
//   * a producer warp streams each visited (page, head) slice -- 4 KB K + 4 KB V (+ 4 x 128 B of per-token
//     scales / zeros for KV4) -- into a shared-memory ring with cp.async.bulk + mbarrier; four compute warps run
//     flash-decoding on 16 tokens of the page each with mma.sync.m16n8k16;
//   * the per-token scale/zero are folded out of the MMAs: nibbles enter as the exact fp16 numbers 1024 + n
//     (even dims) and 1024 + 16 n (odd dims; q is pre-divided by 16 there), one lop3 per pair,
//       q.k   = s_t * (S - 1024 * sum_d q'_d - z_t * sum_d q_d)
//       P.V_d = sum_t (p_t s_t) (1024 + c n) / c  -  1024 * sum_t p_t s_t  -  sum_t p_t s_t z_t
//     so the inner loops never touch a dequant FMA (INT8 pages: 1152 + code, one prmt per pair, per-tensor scales);
//   * split-KV with (max, sum, out) partials merged by the last-arriving CTA (flash-decoding), so the context
//     length is not bounded by shared memory; optional fused per-token INT8 quantisation of the output row.
// Numerics agree with the reference to ~1e-3 of the output scale (north-star tolerance; measured closer to exact
// arithmetic than the reference kernels, tests/test_gpu_parity_r2.py), not bit-exactly; the KV page bytes written for
// the new token follow the reference formula exactly.
#include "kv4_attention.h"
#include "launch.h"
#include "ptx.cuh"

#include <algorithm>
#include <float.h>
#include <mutex>

namespace ob {

constexpr int DH = 128;
constexpr int TPB = 64;            // tokens per page

struct SeqView {
  const int64_t* ktab;  // this sequence's K page pointers
  const int64_t* vtab;
  const int* dyn;       // dynamic page list for this (b, hq) or null
  int mode;             // 0 dense, 1 streaming ring, 2 dynamic page select
  int n_valid;          // cached tokens attended
  int gap, sink_tok, sink_blk, local_blk;
  int rank;             // row of this kv head inside its pool's pages
  int data_bytes;       // H_pool * 64 * 64
  int hpool;
  OB_DEVICE int tab_idx(int pos) const {
    int blk = pos >> 6;
    if (mode == 1) blk = blk < sink_blk ? blk : sink_blk + (blk - sink_blk) % local_blk;
    return blk;
  }
};

// Per-token/per-head asymmetric 4-bit quant of 128 fp16 values held 4 per lane by one warp
// (lane l owns elements 4l..4l+3).  Template.hpp:1063-1081 + Utils.h:1838-1852.
OB_DEVICE void quant_store_token(const __half (&x)[4], uint8_t* page_data_row /*64 B*/, __half* scale_ptr,
                                 __half* zero_ptr, int lane) {
  float mx = -FLT_MAX, mn = FLT_MAX;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float f = __half2float(x[i]);
    mx = fmaxf(mx, f);
    mn = fminf(mn, f);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, m));
  }
  const __half hs = __float2half_rn(__fdividef(mx - mn, 15.0f));
  const __half hz = __float2half_rn(__fdividef(-15.0f * mn, mx - mn));
  const float inv = __fdividef(1.0f, __half2float(hs));
  const float z = __half2float(hz);
  uint32_t q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = f2u8_rni_sat(__fmaf_rn(__half2float(x[i]), inv, z)) & 0xFu;
  const uint16_t packed = (uint16_t)(q[0] | (q[1] << 4) | (q[2] << 8) | (q[3] << 12));
  reinterpret_cast<uint16_t*>(page_data_row)[lane] = packed;
  if (lane == 0) {
    *scale_ptr = hs;
    *zero_ptr = hz;
  }
}

struct AttnParams {
  const __half* q; const __half* k; const __half* v;   // new-token projections
  long long q_bs, k_bs, v_bs;                           // batch strides (elements); head stride = DH
  __half* out;                                          // [B, Hq, DH] contiguous
  const int64_t* r_tab; const int64_t* s_tab;           // [B,2,r_max_pages], [B,2,s_max_pages]
  int r_max_pages, s_max_pages;
  const int* lengths;                                   // [B] incl. the new token
  const int* retrieval_flags; const int* head_rank;     // [Hkv] or null (all retrieval, rank = head)
  const int* dyn_idx; int dyn_pages;                    // [B,Hq,P] or null
  int B, Hq, Hkv;
  int r_hpool, s_hpool;
  int sink_tok, local_tok, sink_blk, local_blk;
  float rope_base, rope_scale; int rotary_dim;
  int n_split;
  float* part_o; float* part_ml; int* counters;         // split-KV workspace
  int timestep;                                         // used when lengths == null
  int sub_chunk, eles_per_ind;                          // LServe page statistics (sparse op): 0 = none
  // fused per-token INT8 quantisation of the attention output (extension): the last CTA of a sequence to finish
  int8_t* q_out; __half* q_scale; __half* q_sum; int* tok_counters;   // quantises the [Hq*128] row; null = off
  // 1 = the caller guarantees that lengths, head tables, page tables, the dynamic page list and every page except the
  // newest were last written BEFORE the kernel that precedes this launch in the stream (the decode loop: the qkv GEMM
  // precedes, pages older than the newest are >= 1 step old), so they may be read while that kernel is still draining.
  // 0 (default of the drop-in ops) = nothing is read before the grid dependency has resolved.
  int stable_history;
  // per-tensor KV8 cache (fused_attention_per_tensor_*): device float[2] = (K, V) dequant / quant scales; null = KV4
  const float* kv_scale_quant_orig; const float* kv_scale_orig_quant;
  long long* dbg_t;   // -DOB_ATT_TIMING builds only (tools/att_timeline.py): [CTA][16] globaltimer stamps
  int dbg_mode;       // -DOB_ATT_TIMING builds only, RESULTS INVALID: 1 = consumers skip the math, 2 = no scale / zero row copies
};

#ifdef OB_ATT_TIMING
OB_DEVICE long long att_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define OB_AT(slot)                                                                                                        \
  do {                                                                                                                     \
    if (p.dbg_t) p.dbg_t[(((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (slot)] = att_gtime(); \
  } while (0)
#else
#define OB_AT(slot) (void)0
#endif
#ifdef OB_ATT_TIMING
#define OB_ATT_MODE(bit) ((p.dbg_mode & (bit)) != 0)
#else
#define OB_ATT_MODE(bit) false
#endif

// invoke_quant(_fuse_sum) (fused_kernels.cu:57-142) of one attention-output row by the first 128 threads of the CTA,
// with the element-to-thread assignment and reduction order of small_ops.cu:quant_kernel (bit-identical results).
OB_DEVICE void quant_row_tail(const AttnParams& p, int b, float* red /*[64] shared*/) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int H = p.Hq * DH, nvec = H >> 3;
  const __half* row = p.out + (size_t)b * H;
  uint4 v[8];
  float amax = 0.f, s = 0.f;
  if (tid < 128) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 128;
      if (idx < nvec) {
        v[i] = __ldcg(reinterpret_cast<const uint4*>(row) + idx);
        const __half* h = reinterpret_cast<const __half*>(&v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float f = __half2float(h[j]);
          s += f;
          amax = fmaxf(amax, fabsf(f));
        }
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, m));
      s += __shfl_xor_sync(0xffffffffu, s, m);
    }
    if (lane == 0) { red[w] = amax; red[32 + w] = s; }
  }
  __syncthreads();
  if (tid < 128) {
    float x = lane < 4 ? red[lane] : -3.0e38f;
    float y = lane < 4 ? red[32 + lane] : 0.f;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, m));
      y += __shfl_xor_sync(0xffffffffu, y, m);
    }
    if (tid == 0) {
      p.q_scale[b] = __float2half_rn(__fdividef(x, 127.0f));
      if (p.q_sum) p.q_sum[b] = __float2half_rn(y);
    }
    const float qs = __fdividef(127.0f, x);
    uint2* dst = reinterpret_cast<uint2*>(p.q_out + (size_t)b * H);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 128;
      if (idx < nvec) {
        const __half* h = reinterpret_cast<const __half*>(&v[i]);
        uint32_t bq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bq[j] = (uint32_t)(uint8_t)f2i8_rni_sat(__half2float(h[j]) * qs);
        uint2 r;
        r.x = bq[0] | (bq[1] << 8) | (bq[2] << 16) | (bq[3] << 24);
        r.y = bq[4] | (bq[5] << 8) | (bq[6] << 16) | (bq[7] << 24);
        dst[idx] = r;
      }
    }
  }
}

// Called by every thread of a CTA that has just written final outputs of sequence b.
OB_DEVICE void fused_quant_tail(const AttnParams& p, int b, int* flag, float* red) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *flag = (atomicAdd(&p.tok_counters[b], 1) == (int)gridDim.y - 1);
  __syncthreads();
  if (!*flag) return;
  __threadfence();
  quant_row_tail(p, b, red);
  if (threadIdx.x == 0) p.tok_counters[b] = 0;
}

// ------------------------------------------------------------------------------------------------
// Decode kernel.  160 threads: warps 0-3 compute, warp 4 = bulk-copy producer.
//   * every visited (page, kv-head) slice -- 4 KB of K nibbles, 4 KB of V nibbles, 4 x 128 B of fp16
//     scales / zeros -- is contiguous in the reference page layout, so the producer streams it into a
//     4-stage shared-memory ring with cp.async.bulk + mbarrier (no per-thread global loads);
//   * each compute warp owns 16 tokens of the page and runs flash-decoding on them with
//     mma.sync.m16n8k16 (f16 x f16 -> f32):  S^T[16 tok x 8 heads] = (K nibbles - 8) . Q^T, online softmax,
//     P^T moved into B-fragment form with movmatrix, O^T[128 dims x 8 heads] += (V nibbles - 8)^T . P'^T.
//     The nibbles enter the MMAs as exact small integers; scale / zero are folded outside the MMAs (see
//     header comment).  Token <-> MMA-row assignment kappa() is chosen so that both the K (LDS.128) and V
//     (LDS.64) fragment loads are at most 2-way bank conflicted.
//   * the four warps (and, for the split that owns it, the new token) are merged like KV splits.
// ------------------------------------------------------------------------------------------------
#ifndef OB_ATT_PAIR
#define OB_ATT_PAIR 1       // KV4 page loop: two pages per round (0 = the one-page loop, kept for A/B and used by KV8)
#endif
constexpr int V2_STAGES = 4;
constexpr int V2_STAGE_BYTES = 2 * 4096 + 4 * 128;
// per-tensor KV8 pages (fused_attention_per_tensor_*, cache_engine.py:73-88): a (page, head) slice is 64 tokens x 128 int8 of
// K and of V, no per-token scale rows
constexpr int V8_STAGES = 3;
constexpr int V8_STAGE_BYTES = 2 * 8192;
constexpr int V2_THREADS = 160;
constexpr float LOG2E = 1.4426950408889634f;

OB_DEVICE int kappa(int r) { return (r & 8) | ((r & 7) >> 1) | ((r & 1) << 2); }

OB_DEVICE void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
OB_DEVICE uint32_t movmatrix_trans(uint32_t x) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(x));
  return d;
}
OB_DEVICE uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
OB_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Visit { int tab, lo, hi; };

struct Visits {
  int mode, n, tl;
  int a1, b0, nA;       // streaming: range A = [0,a1), range B = [b0,tl); nA pages in A
  int sink_blk, local_blk, dyn_pages;
  const int* dyn;
  OB_DEVICE void init(const SeqView& sv, int tl_, int dyn_pages_) {
    mode = sv.mode; tl = tl_; dyn = sv.dyn; dyn_pages = dyn_pages_;
    sink_blk = sv.sink_blk; local_blk = sv.local_blk;
    if (mode == 0) n = (tl + TPB - 1) / TPB;
    else if (mode == 2) n = tl > 0 ? dyn_pages : 0;
    else {
      a1 = min(sv.sink_tok, sv.n_valid);
      b0 = sv.n_valid > sv.sink_tok ? sv.sink_tok + sv.gap : tl;
      nA = (a1 + TPB - 1) / TPB;
      const int nB = b0 < tl ? ((tl - 1) >> 6) - (b0 >> 6) + 1 : 0;
      n = nA + nB;
    }
  }
  OB_DEVICE Visit get(int v) const {
    Visit r;
    if (mode == 0) {
      r.tab = v; r.lo = 0; r.hi = min(TPB, tl - v * TPB);
    } else if (mode == 2) {
      r.tab = dyn[v]; r.lo = 0; r.hi = (v == dyn_pages - 1) ? ((tl - 1) % TPB + 1) : TPB;
    } else {
      int x0, x1, blk;
      if (v < nA) { x0 = 0; x1 = a1; blk = v; }
      else { x0 = b0; x1 = tl; blk = (b0 >> 6) + (v - nA); }
      r.lo = max(x0, blk * TPB) - blk * TPB;
      r.hi = min(x1, blk * TPB + TPB) - blk * TPB;
      r.tab = blk < sink_blk ? blk : sink_blk + (blk - sink_blk) % local_blk;
    }
    return r;
  }
};

template <bool KV8>
__global__ void __launch_bounds__(V2_THREADS, KV8 ? 3 : 4)
kv4_decode_kernel(const AttnParams p, const int G) {
  constexpr int NSTAGE = KV8 ? V8_STAGES : V2_STAGES;
  constexpr int STAGE_BYTES = KV8 ? V8_STAGE_BYTES : V2_STAGE_BYTES;
  constexpr int ROW_BYTES = KV8 ? DH : DH / 2;          // bytes of one cached token of one head
  constexpr int SLICE = TPB * ROW_BYTES;                 // bytes of one (page, head) slice of K or of V
  extern __shared__ __align__(128) uint8_t ring[];  // NSTAGE * STAGE_BYTES, reused for the final merge
  __shared__ __align__(16) __half q_s[8][DH];
  __shared__ __align__(16) __half kv_new[2][DH];
  __shared__ float qsum_s[8], qbias_s[8], cur_logit_s[8];
  __shared__ float rope_cs[DH / 2], rope_sn[DH / 2];
  __shared__ int64_t kptr_s[32], vptr_s[32];
  __shared__ float ml_s[4][8][2];
  __shared__ __align__(8) uint64_t full[NSTAGE], empty[NSTAGE];
  __shared__ int flag_s;
  __shared__ float red_s[64];

  pdl_trigger();
  if (threadIdx.x == 0) OB_AT(0);      // entry
  if (!p.stable_history) pdl_wait();   // e.g. a decode launched straight after the prefill writer of the same pages
  const int split = blockIdx.x;
  const int b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int group = p.Hq / p.Hkv;
  const int hq0 = (G == 1 && group != 1) ? blockIdx.y : blockIdx.y * G;
  const int hkv = hq0 / group;
  const bool is_retrieval = p.retrieval_flags ? (p.retrieval_flags[hkv] != 0) : true;
  const int tl = (p.lengths ? p.lengths[b] : p.timestep + 1) - 1;  // cached tokens == position of the new one

  SeqView sv;
  sv.rank = p.head_rank ? p.head_rank[hkv] : hkv;
  if (is_retrieval) {
    sv.ktab = p.r_tab + (size_t)b * 2 * p.r_max_pages;
    sv.vtab = sv.ktab + p.r_max_pages;
    sv.hpool = p.r_hpool;
    sv.mode = 0;
    sv.n_valid = tl;
    sv.dyn = nullptr;
    if (p.dyn_idx) {
      sv.mode = 2;
      sv.dyn = p.dyn_idx + ((size_t)b * p.Hq + hq0) * p.dyn_pages;
    }
    sv.gap = 0; sv.sink_tok = 0; sv.sink_blk = 0; sv.local_blk = 1;
  } else {
    sv.ktab = p.s_tab + (size_t)b * 2 * p.s_max_pages;
    sv.vtab = sv.ktab + p.s_max_pages;
    sv.hpool = p.s_hpool;
    sv.mode = 1;
    sv.dyn = nullptr;
    sv.n_valid = min(p.sink_tok + p.local_tok - 1, tl);
    sv.gap = tl - sv.n_valid;
    sv.sink_tok = p.sink_tok; sv.sink_blk = p.sink_blk; sv.local_blk = p.local_blk;
  }
  sv.data_bytes = sv.hpool * SLICE;
  Visits vis;
  vis.init(sv, tl, p.dyn_pages);
  const int per_split = (vis.n + p.n_split - 1) / p.n_split;
  const int v0 = min(split * per_split, vis.n);
  const int v1 = min(v0 + per_split, vis.n);
  const bool owns_current = (split == p.n_split - 1);

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 4); }
    mbar_fence_init();
  }
  __syncthreads();

  const float qk_scale = 0.08838834764831845f * LOG2E;  // 1/sqrt(128) * log2(e)
  float acc[8][4];
  float m0 = -1.0e30f, m1 = -1.0e30f, l0 = 0.f, l1 = 0.f, corr0 = 0.f, corr1 = 0.f, sp0 = 0.f, sp1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

  if (warp == 4) {
    // ================================================================ producer (starts before the prologue)
    // Page pointers are fetched 32 visits at a time by the whole warp into shared memory so that the single
    // issuing lane never waits on a dependent global load between bulk copies.
    int s = 0, ph = 0;
    for (int vb = v0; vb < v1; vb += 32) {
      const int nv = min(32, v1 - vb);
      if (lane < nv) {
        const Visit vv = vis.get(vb + lane);
        kptr_s[lane] = sv.ktab[vv.tab];
        vptr_s[lane] = sv.vtab[vv.tab];
      }
      __syncwarp();
      if (lane == 0) {
        for (int i = 0; i < nv; ++i) {
          // Pages older than the newest one were written at least two decode steps ago and are streamed while the
          // previous kernel (the qkv GEMM) is still draining; the newest page holds the token appended by the
          // previous step's attention call, so it is fetched only after the grid dependency resolved.
          if (vb + i == vis.n - 1) pdl_wait();
          mbar_wait(&empty[s], ph ^ 1);
          const uint8_t* kp = reinterpret_cast<const uint8_t*>(kptr_s[i]);
          const uint8_t* vp = reinterpret_cast<const uint8_t*>(vptr_s[i]);
          uint8_t* st = ring + s * STAGE_BYTES;
          if (vb + i == v0) OB_AT(8);           // producer: first page issued
          if (vb + i == v1 - 1) OB_AT(9);       // producer: last page issued
          mbar_arrive_expect_tx(&full[s], OB_ATT_MODE(2) ? 2 * SLICE : STAGE_BYTES);
          bulk_g2s(st, kp + (size_t)sv.rank * SLICE, SLICE, &full[s]);
          bulk_g2s(st + SLICE, vp + (size_t)sv.rank * SLICE, SLICE, &full[s]);
          if (!KV8 && !OB_ATT_MODE(2)) {
            bulk_g2s(st + 8192, kp + sv.data_bytes + sv.rank * 128, 128, &full[s]);
            bulk_g2s(st + 8192 + 128, kp + sv.data_bytes + (sv.hpool + sv.rank) * 128, 128, &full[s]);
            bulk_g2s(st + 8192 + 256, vp + sv.data_bytes + sv.rank * 128, 128, &full[s]);
            bulk_g2s(st + 8192 + 384, vp + sv.data_bytes + (sv.hpool + sv.rank) * 128, 128, &full[s]);
          }
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ prologue: q/k RoPE, append new K/V
    // cos/sin depend on (position, pair index) only: 64 accurate powf/sincosf per CTA, shared by all heads, and
    // evaluated before the grid dependency resolves (positions were written many kernels ago).
    {
      const int half_rot = p.rotary_dim >> 1;
      if (tid < DH / 2 && tid < half_rot) {
        const float inv_freq = ((float)tl * p.rope_scale) / powf(p.rope_base, (float)(2 * tid) / (float)p.rotary_dim);
        sincosf(inv_freq, &rope_sn[tid], &rope_cs[tid]);
      }
    }
    if (tid == 0) OB_AT(1);   // pre-dependency prologue done (tables, RoPE angles)
    pdl_wait();  // q, k, v are the previous kernel's output
    if (tid == 0) OB_AT(2);   // grid dependency resolved
    asm volatile("bar.sync 1, 128;" ::: "memory");
    {
      const int half_rot = p.rotary_dim >> 1;
      for (int item = tid; item < 9 * (DH / 2); item += 128) {
        const int h = item / (DH / 2), d = item - h * (DH / 2);
        if (h < 8 && h >= G) {  // unused head rows of the 8-wide MMA N dimension
          q_s[h][d] = __float2half_rn(0.f);
          q_s[h][d + DH / 2] = __float2half_rn(0.f);
          continue;
        }
        const __half* src = (h < 8) ? p.q + (size_t)b * p.q_bs + (size_t)(hq0 + h) * DH
                                    : p.k + (size_t)b * p.k_bs + (size_t)hkv * DH;
        __half* dst = (h < 8) ? q_s[h] : kv_new[0];
        if (d < half_rot) {
          const float sn = rope_sn[d], cs = rope_cs[d];
          const float x = __half2float(src[d]), y = __half2float(src[d + half_rot]);
          dst[d] = __float2half_rn(cs * x - sn * y);
          dst[d + half_rot] = __float2half_rn(cs * y + sn * x);
        } else {
          const int e = p.rotary_dim + 2 * (d - half_rot);
          dst[e] = src[e];
          dst[e + 1] = src[e + 1];
        }
      }
      for (int d = tid; d < DH; d += 128) kv_new[1][d] = p.v[(size_t)b * p.v_bs + (size_t)hkv * DH + d];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (tid < 8) { qsum_s[tid] = 0.f; qbias_s[tid] = 0.f; cur_logit_s[tid] = -1.0e30f; }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int h = warp; h < G; h += 4) {
      // Full-precision logit of the new token (Template.hpp:1356-1376, log2 domain), then re-encode q for the
      // biased-nibble MMAs: odd dims are stored as q/16 because their nibbles enter as 1024 + 16 n.
      float s = 0.f, dot = 0.f, bias = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = lane * 4 + i;
        const float qv = __half2float(q_s[h][d]);
        dot += qv * __half2float(kv_new[0][d]);
        if (KV8) {
          bias += qv;          // int8 codes enter the MMAs as 1152 + code (1024 + (code ^ 0x80)): no per-dimension rescale
        } else if (i & 1) {
          const __half qh = __float2half_rn(qv * 0.0625f);
          q_s[h][d] = qh;
          const float qe = __half2float(qh);
          bias += qe;
          s += 16.f * qe;
        } else {
          bias += qv;
          s += qv;
        }
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, m);
        dot += __shfl_xor_sync(0xffffffffu, dot, m);
        bias += __shfl_xor_sync(0xffffffffu, bias, m);
      }
      if (lane == 0) { qsum_s[h] = s; qbias_s[h] = (KV8 ? 1152.f : 1024.f) * bias; cur_logit_s[h] = dot * qk_scale; }
    }
    const bool writer = owns_current && ((G > 1) || group == 1 || (hq0 == hkv * group));
    if (writer && warp >= 2) {
      const int which = warp - 2;  // 0 = K, 1 = V
      const int64_t* tab = which ? sv.vtab : sv.ktab;
      SeqView wv = sv;
      if (wv.mode == 2) wv.mode = 0;
      uint8_t* page = reinterpret_cast<uint8_t*>(tab[wv.tab_idx(tl)]);
      const int slot = tl & 63;
      __half x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = kv_new[which][lane * 4 + i];
      if (KV8) {
        // per-tensor INT8: code = cvt.rni.sat.s8(x * scale_orig_quant) (common/...Utils.h:2041-2048); 128 B per token row
        const float sq = p.kv_scale_orig_quant[which];
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) w |= ((uint32_t)(uint8_t)f2i8_rni_sat(__half2float(x[i]) * sq)) << (8 * i);
        reinterpret_cast<uint32_t*>(page + (size_t)sv.rank * SLICE + slot * ROW_BYTES)[lane] = w;
      } else {
        __half* sc = reinterpret_cast<__half*>(page + sv.data_bytes) + sv.rank * TPB + slot;
        quant_store_token(x, page + (size_t)sv.rank * TPB * (DH / 2) + slot * (DH / 2), sc, sc + sv.hpool * TPB, lane);
      }
      if (which == 0 && is_retrieval && p.sub_chunk > 0) {
        // LServe: fold the appended post-RoPE key into the kmax / kmin statistics of its sub-chunk, element-wise
        // against what the page holds (sparse_attention/...Template.hpp:1414-1429; fmaxf / fminf on fp16 values).
        __half* stats = reinterpret_cast<__half*>(page + sv.data_bytes) + 2 * sv.hpool * TPB;
        __half* mxp = stats + (size_t)(slot / p.sub_chunk) * p.eles_per_ind + sv.rank * DH + lane * 4;
        __half* mnp = mxp + (size_t)(TPB / p.sub_chunk) * p.eles_per_ind;
        uint2 a = *reinterpret_cast<const uint2*>(mxp), c = *reinterpret_cast<const uint2*>(mnp);
        __half* ah = reinterpret_cast<__half*>(&a);
        __half* ch = reinterpret_cast<__half*>(&c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ah[i] = __float2half_rn(fmaxf(__half2float(ah[i]), __half2float(x[i])));
          ch[i] = __float2half_rn(fminf(__half2float(ch[i]), __half2float(x[i])));
        }
        *reinterpret_cast<uint2*>(mxp) = a;
        *reinterpret_cast<uint2*>(mnp) = c;
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");

    // ================================================================ compute warps
    if (tid == 0) OB_AT(3);   // q / k RoPE, new-token logit, append done: page loop starts
    const int g = lane >> 2, c = lane & 3;
    uint32_t qB[16];
    if (KV8) {
      // K-step s (16 dims) of the S^T MMAs: k-index 2c+i <-> dim 32c+4s+i, k-index 2c+8+i <-> dim 32c+4s+2+i (i = 0, 1): the four
      // bytes a lane reads per token row and K-step are one 32-bit word (any bijection works as long as Q uses the same one)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        qB[2 * ks] = h2_as_u32(__halves2half2(q_s[g][c * 32 + 4 * ks], q_s[g][c * 32 + 4 * ks + 1]));
        qB[2 * ks + 1] = h2_as_u32(__halves2half2(q_s[g][c * 32 + 4 * ks + 2], q_s[g][c * 32 + 4 * ks + 3]));
      }
    } else {
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          qB[w * 4 + j] = h2_as_u32(__halves2half2(q_s[g][c * 32 + w * 8 + j], q_s[g][c * 32 + w * 8 + j + 4]));
    }
    const float qs0 = qsum_s[2 * c], qs1 = qsum_s[2 * c + 1];
    const float qb0 = qbias_s[2 * c], qb1 = qbias_s[2 * c + 1];
    const float k_ts = KV8 ? p.kv_scale_quant_orig[0] : 1.f;   // per-tensor dequant scale of K
    const int base = warp * 16;
    const int tok_a = base + kappa(g), tok_b = base + kappa(g + 8);
    const uint32_t ka_off = tok_a * ROW_BYTES + c * (ROW_BYTES / 4), kb_off = tok_b * ROW_BYTES + c * (ROW_BYTES / 4);
    constexpr int VCH = ROW_BYTES / 8;     // bytes of the 16 dims [16g, 16g+16) of one token: 8 (nibbles) / 16 (int8)
    const uint32_t v_off0 = SLICE + (base + kappa(2 * c)) * ROW_BYTES + g * VCH, v_off1 = SLICE + (base + kappa(2 * c + 1)) * ROW_BYTES + g * VCH;
    const uint32_t v_off2 = SLICE + (base + kappa(8 + 2 * c)) * ROW_BYTES + g * VCH, v_off3 = SLICE + (base + kappa(9 + 2 * c)) * ROW_BYTES + g * VCH;
    int s = 0, ph = 0;
#if OB_ATT_PAIR
    if constexpr (!KV8) {
      // ---------------------------------------------------------------------------------------------------------------
      // KV4 page loop, TWO pages per round.  The page loop is bound by the math of the four compute warps, not by the loads
      // (profiles/r2_att_timeline.log: with the math switched off the same loop streams at 6.8 TB/s, with it 3.2 TB/s; ~8
      // stall cycles per issued instruction with 3.5 warps per scheduler).  Two pages per round give every warp two
      // independent K chains, one max / rescale / branch sequence per 32 tokens instead of 16, and two MMAs per accumulator
      // in the V phase; shared-memory and barrier addresses are 32-bit values formed once, masks are selects.
      // An odd tail page runs as a pair whose second half is masked (its stage pointer aliases the first).
      // ---------------------------------------------------------------------------------------------------------------
      const uint32_t ring_u32 = smem_u32(ring), full_u32 = smem_u32(full), empty_u32 = smem_u32(empty);
      auto k_phase = [&](uint32_t st, float (&sa)[4], float (&sb)[4]) {
        const uint4 ka = lds_v4(st + ka_off), kb = lds_v4(st + kb_off);
        const uint32_t kaw[4] = {ka.x, ka.y, ka.z, ka.w}, kbw[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
          const uint32_t ta = kaw[w] >> 8, tb = kbw[w] >> 8;
          asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(a0) : "r"(kaw[w]));
          asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(a1) : "r"(kaw[w]));
          asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(a2) : "r"(ta));
          asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(a3) : "r"(ta));
          asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(b0) : "r"(kbw[w]));
          asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(b1) : "r"(kbw[w]));
          asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(b2) : "r"(tb));
          asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(b3) : "r"(tb));
          mma16816(sa, a0, b0, a1, b1, qB[4 * w], qB[4 * w + 1]);
          mma16816(sb, a2, b2, a3, b3, qB[4 * w + 2], qB[4 * w + 3]);
        }
      };
      // log2-domain logits of this lane's two tokens (a, b) x two heads (2c, 2c+1) of one page; masked tokens -> -inf
      auto logits = [&](uint32_t st, const float (&sa)[4], const float (&sb)[4], bool va, bool vb, float (&lg)[4]) {
        const float ksa = lds_f16(st + 8192 + 2 * tok_a), kza = lds_f16(st + 8192 + 128 + 2 * tok_a);
        const float ksb = lds_f16(st + 8192 + 2 * tok_b), kzb = lds_f16(st + 8192 + 128 + 2 * tok_b);
        const float a0 = ksa * ((sa[0] + sb[0]) - qb0 - kza * qs0) * qk_scale, a1 = ksa * ((sa[1] + sb[1]) - qb1 - kza * qs1) * qk_scale;
        const float b0 = ksb * ((sa[2] + sb[2]) - qb0 - kzb * qs0) * qk_scale, b1 = ksb * ((sa[3] + sb[3]) - qb1 - kzb * qs1) * qk_scale;
        lg[0] = va ? a0 : -INFINITY; lg[1] = va ? a1 : -INFINITY;
        lg[2] = vb ? b0 : -INFINITY; lg[3] = vb ? b1 : -INFINITY;
      };
      // probabilities of one page -> P'^T B fragments (scaled by the V scale), softmax sums and zero-point corrections
      auto probs = [&](uint32_t st, const float (&lg)[4], bool va, bool vb, uint32_t& pb_lo, uint32_t& pb_hi) {
        const float pa0 = ex2(lg[0] - m0), pa1 = ex2(lg[1] - m1), pb0 = ex2(lg[2] - m0), pb1 = ex2(lg[3] - m1);
        l0 += pa0 + pb0;
        l1 += pa1 + pb1;
        const float vsa_ = lds_f16(st + 8192 + 256 + 2 * tok_a), vza_ = lds_f16(st + 8192 + 384 + 2 * tok_a);
        const float vsb_ = lds_f16(st + 8192 + 256 + 2 * tok_b), vzb_ = lds_f16(st + 8192 + 384 + 2 * tok_b);
        const float vsa = va ? vsa_ : 0.f, vza = va ? vza_ : 0.f, vsb = vb ? vsb_ : 0.f, vzb = vb ? vzb_ : 0.f;
        const __half2 ha = __floats2half2_rn(pa0 * vsa, pa1 * vsa), hb = __floats2half2_rn(pb0 * vsb, pb1 * vsb);
        const float2 fa = __half22float2(ha), fb = __half22float2(hb);  // the values the MMA will really use
        sp0 += fa.x + fb.x;
        sp1 += fa.y + fb.y;
        corr0 += fa.x * vza + fb.x * vzb;
        corr1 += fa.y * vza + fb.y * vzb;
        pb_lo = movmatrix_trans(h2_as_u32(ha));
        pb_hi = movmatrix_trans(h2_as_u32(hb));
      };
      auto v_phase = [&](uint32_t st, uint32_t pb_lo, uint32_t pb_hi) {
        const uint2 w0 = lds_v2(st + v_off0), w1 = lds_v2(st + v_off1), w2 = lds_v2(st + v_off2), w3 = lds_v2(st + v_off3);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t A = j < 4 ? w0.x : w0.y, B = j < 4 ? w1.x : w1.y, Cw = j < 4 ? w2.x : w2.y, D = j < 4 ? w3.x : w3.y;
          const uint32_t sel = (uint32_t)(j & 3) | ((uint32_t)(4 + (j & 3)) << 8);
          const uint32_t u01 = __byte_perm(A, B, sel), u23 = __byte_perm(Cw, D, sel);
          uint32_t t0, t1, t2, t3;
          asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(t0) : "r"(u01));
          asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(t1) : "r"(u01));
          asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(t2) : "r"(u23));
          asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(t3) : "r"(u23));
          mma16816(acc[j], t0, t1, t2, t3, pb_lo, pb_hi);
        }
      };
      for (int v = v0; v < v1; v += 2) {
        const bool two = v + 1 < v1;
        const Visit x0 = vis.get(v);
        const Visit x1 = two ? vis.get(v + 1) : x0;
        const int s1 = two ? (s + 1 == NSTAGE ? 0 : s + 1) : s;
        const int ph1 = (two && s + 1 == NSTAGE) ? ph ^ 1 : ph;
        mbar_wait_a(full_u32 + s * 8, ph);
        if (two) mbar_wait_a(full_u32 + s1 * 8, ph1);
        const uint32_t st0 = ring_u32 + s * STAGE_BYTES, st1 = ring_u32 + s1 * STAGE_BYTES;
        float sa0[4] = {0.f, 0.f, 0.f, 0.f}, sb0[4] = {0.f, 0.f, 0.f, 0.f}, sa1[4] = {0.f, 0.f, 0.f, 0.f}, sb1[4] = {0.f, 0.f, 0.f, 0.f};
        k_phase(st0, sa0, sb0);
        k_phase(st1, sa1, sb1);
        const bool va0 = tok_a >= x0.lo && tok_a < x0.hi, vb0 = tok_b >= x0.lo && tok_b < x0.hi;
        const bool va1 = two && tok_a >= x1.lo && tok_a < x1.hi, vb1 = two && tok_b >= x1.lo && tok_b < x1.hi;
        float lg0[4], lg1[4];
        logits(st0, sa0, sb0, va0, vb0, lg0);
        logits(st1, sa1, sb1, va1, vb1, lg1);
        // ---------------- online softmax over the 32 tokens (lanes with equal c share the heads 2c, 2c+1)
        float x0m = fmaxf(fmaxf(lg0[0], lg0[2]), fmaxf(lg1[0], lg1[2])), x1m = fmaxf(fmaxf(lg0[1], lg0[3]), fmaxf(lg1[1], lg1[3]));
#pragma unroll
        for (int k = 4; k <= 16; k <<= 1) {
          x0m = fmaxf(x0m, __shfl_xor_sync(0xffffffffu, x0m, k));
          x1m = fmaxf(x1m, __shfl_xor_sync(0xffffffffu, x1m, k));
        }
        const float n0 = fmaxf(m0, x0m), n1 = fmaxf(m1, x1m);
        if (__any_sync(0xffffffffu, n0 != m0 || n1 != m1)) {
          const float f0 = ex2(m0 - n0), f1 = ex2(m1 - n1);
          l0 *= f0; l1 *= f1; corr0 *= f0; corr1 *= f1; sp0 *= f0; sp1 *= f1;
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc[j][0] *= f0; acc[j][1] *= f1; acc[j][2] *= f0; acc[j][3] *= f1; }
          m0 = n0; m1 = n1;
        }
        uint32_t p0_lo, p0_hi, p1_lo, p1_hi;
        probs(st0, lg0, va0, vb0, p0_lo, p0_hi);
        probs(st1, lg1, va1, vb1, p1_lo, p1_hi);
        v_phase(st0, p0_lo, p0_hi);
        v_phase(st1, p1_lo, p1_hi);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_a(empty_u32 + s * 8);
          if (two) mbar_arrive_a(empty_u32 + s1 * 8);
        }
        if (++s == NSTAGE) { s = 0; ph ^= 1; }
        if (two && ++s == NSTAGE) { s = 0; ph ^= 1; }
      }
    } else
#endif
    for (int v = v0; v < v1; ++v) {
      const Visit vv = vis.get(v);
      mbar_wait(&full[s], ph);
      if (base < vv.hi && base + 16 > vv.lo && !OB_ATT_MODE(1)) {
        const uint8_t* st = ring + s * STAGE_BYTES;
        const __half* ksc = reinterpret_cast<const __half*>(st + 8192);
        float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
        uint4 W0, W1, W2, W3;              // V bytes of this lane's four tokens (KV8: 16 B each; KV4: the low 8 B are used)
        if (KV8) {
          // ---------------- S^T = (1152 + k8) . Q^T   (bias removed below); int8 -> exact fp16 1024 + (code ^ 0x80) with one prmt
          const uint4 ka0 = *reinterpret_cast<const uint4*>(st + ka_off), ka1 = *reinterpret_cast<const uint4*>(st + ka_off + 16);
          const uint4 kb0 = *reinterpret_cast<const uint4*>(st + kb_off), kb1 = *reinterpret_cast<const uint4*>(st + kb_off + 16);
          W0 = *reinterpret_cast<const uint4*>(st + v_off0);
          W1 = *reinterpret_cast<const uint4*>(st + v_off1);
          W2 = *reinterpret_cast<const uint4*>(st + v_off2);
          W3 = *reinterpret_cast<const uint4*>(st + v_off3);
          const uint32_t kaw[8] = {ka0.x, ka0.y, ka0.z, ka0.w, ka1.x, ka1.y, ka1.z, ka1.w};
          const uint32_t kbw[8] = {kb0.x, kb0.y, kb0.z, kb0.w, kb1.x, kb1.y, kb1.z, kb1.w};
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t xa = kaw[ks] ^ 0x80808080u, xb = kbw[ks] ^ 0x80808080u;
            const uint32_t a0 = __byte_perm(xa, 0x64646464u, 0x5140), a2 = __byte_perm(xa, 0x64646464u, 0x7362);
            const uint32_t a1 = __byte_perm(xb, 0x64646464u, 0x5140), a3 = __byte_perm(xb, 0x64646464u, 0x7362);
            if (ks & 1) mma16816(sb, a0, a1, a2, a3, qB[2 * ks], qB[2 * ks + 1]);
            else mma16816(sa, a0, a1, a2, a3, qB[2 * ks], qB[2 * ks + 1]);
          }
        } else {
        // ---------------- S^T = (1024 + c n_K) . Q'^T     (bias removed below)
        const uint4 ka = *reinterpret_cast<const uint4*>(st + ka_off);
        const uint4 kb = *reinterpret_cast<const uint4*>(st + kb_off);
        const uint2 w0 = *reinterpret_cast<const uint2*>(st + v_off0);
        const uint2 w1 = *reinterpret_cast<const uint2*>(st + v_off1);
        const uint2 w2 = *reinterpret_cast<const uint2*>(st + v_off2);
        const uint2 w3 = *reinterpret_cast<const uint2*>(st + v_off3);
        W0 = make_uint4(w0.x, w0.y, 0, 0); W1 = make_uint4(w1.x, w1.y, 0, 0);
        W2 = make_uint4(w2.x, w2.y, 0, 0); W3 = make_uint4(w3.x, w3.y, 0, 0);
        const uint32_t kaw[4] = {ka.x, ka.y, ka.z, ka.w}, kbw[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
          const uint32_t ta = kaw[w] >> 8, tb = kbw[w] >> 8;
          asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(a0) : "r"(kaw[w]));
          asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(a1) : "r"(kaw[w]));
          asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(a2) : "r"(ta));
          asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(a3) : "r"(ta));
          asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(b0) : "r"(kbw[w]));
          asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(b1) : "r"(kbw[w]));
          asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(b2) : "r"(tb));
          asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(b3) : "r"(tb));
          // slabs 2w (pairs j = 0,1) and 2w+1 (pairs j = 2,3); two independent accumulator chains
          mma16816(sa, a0, b0, a1, b1, qB[4 * w], qB[4 * w + 1]);
          mma16816(sb, a2, b2, a3, b3, qB[4 * w + 2], qB[4 * w + 3]);
        }
        }
        const bool full_page = (vv.lo == 0) & (vv.hi == TPB);
        const bool va = full_page || (tok_a >= vv.lo && tok_a < vv.hi);
        const bool vb = full_page || (tok_b >= vv.lo && tok_b < vv.hi);
        float la0, la1, lb0, lb1;
        if (KV8) {
          la0 = va ? k_ts * ((sa[0] + sb[0]) - qb0) * qk_scale : -INFINITY;
          la1 = va ? k_ts * ((sa[1] + sb[1]) - qb1) * qk_scale : -INFINITY;
          lb0 = vb ? k_ts * ((sa[2] + sb[2]) - qb0) * qk_scale : -INFINITY;
          lb1 = vb ? k_ts * ((sa[3] + sb[3]) - qb1) * qk_scale : -INFINITY;
        } else {
          const float ksa = __half2float(ksc[tok_a]), kza = __half2float(ksc[64 + tok_a]);
          const float ksb = __half2float(ksc[tok_b]), kzb = __half2float(ksc[64 + tok_b]);
          la0 = va ? ksa * ((sa[0] + sb[0]) - qb0 - kza * qs0) * qk_scale : -INFINITY;
          la1 = va ? ksa * ((sa[1] + sb[1]) - qb1 - kza * qs1) * qk_scale : -INFINITY;
          lb0 = vb ? ksb * ((sa[2] + sb[2]) - qb0 - kzb * qs0) * qk_scale : -INFINITY;
          lb1 = vb ? ksb * ((sa[3] + sb[3]) - qb1 - kzb * qs1) * qk_scale : -INFINITY;
        }
        // ---------------- online softmax over the 16 tokens (lanes with equal c share the heads 2c, 2c+1)
        float x0 = fmaxf(la0, lb0), x1 = fmaxf(la1, lb1);
#pragma unroll
        for (int k = 4; k <= 16; k <<= 1) {
          x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, k));
          x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, k));
        }
        const float n0 = fmaxf(m0, x0), n1 = fmaxf(m1, x1);
        if (__any_sync(0xffffffffu, n0 != m0 || n1 != m1)) {
          const float f0 = ex2(m0 - n0), f1 = ex2(m1 - n1);
          l0 *= f0; l1 *= f1; corr0 *= f0; corr1 *= f1; sp0 *= f0; sp1 *= f1;
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc[j][0] *= f0; acc[j][1] *= f1; acc[j][2] *= f0; acc[j][3] *= f1; }
          m0 = n0; m1 = n1;
        }
        const float pa0 = ex2(la0 - m0), pa1 = ex2(la1 - m1), pb0 = ex2(lb0 - m0), pb1 = ex2(lb1 - m1);
        l0 += pa0 + pb0;
        l1 += pa1 + pb1;
        float vsa, vza, vsb, vzb;
        if (KV8) {   // per-tensor V scale is applied once at the end; masked tokens contribute nothing
          vsa = va ? 1.f : 0.f; vsb = vb ? 1.f : 0.f; vza = 0.f; vzb = 0.f;
        } else {
          vsa = va ? __half2float(ksc[128 + tok_a]) : 0.f; vza = va ? __half2float(ksc[192 + tok_a]) : 0.f;
          vsb = vb ? __half2float(ksc[128 + tok_b]) : 0.f; vzb = vb ? __half2float(ksc[192 + tok_b]) : 0.f;
        }
        const __half2 ha = __floats2half2_rn(pa0 * vsa, pa1 * vsa), hb = __floats2half2_rn(pb0 * vsb, pb1 * vsb);
        const float2 fa = __half22float2(ha), fb = __half22float2(hb);  // the values the MMA will really use
        sp0 += fa.x + fb.x;
        sp1 += fa.y + fb.y;
        corr0 += fa.x * vza + fb.x * vzb;
        corr1 += fa.y * vza + fb.y * vzb;
        const uint32_t pb_lo = movmatrix_trans(h2_as_u32(ha));
        const uint32_t pb_hi = movmatrix_trans(h2_as_u32(hb));
        if (KV8) {
          // ---------------- O^T += (1152 + v8)^T . P^T : MMA j, row g <-> dim 16g+2j, row g+8 <-> dim 16g+2j+1; k <-> this lane's 4 tokens
          const uint32_t w0a[4] = {W0.x, W0.y, W0.z, W0.w}, w1a[4] = {W1.x, W1.y, W1.z, W1.w};
          const uint32_t w2a[4] = {W2.x, W2.y, W2.z, W2.w}, w3a[4] = {W3.x, W3.y, W3.z, W3.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t sel = (j & 1) ? 0x7362u : 0x5140u;   // bytes (2, 6, 3, 7) or (0, 4, 1, 5): dims 2j, 2j+1 of both tokens
            const uint32_t g01 = __byte_perm(w0a[j >> 1], w1a[j >> 1], sel) ^ 0x80808080u;   // [t0 d, t1 d, t0 d+1, t1 d+1]
            const uint32_t g23 = __byte_perm(w2a[j >> 1], w3a[j >> 1], sel) ^ 0x80808080u;
            const uint32_t t0 = __byte_perm(g01, 0x64646464u, 0x5140), t1 = __byte_perm(g01, 0x64646464u, 0x7362);
            const uint32_t t2 = __byte_perm(g23, 0x64646464u, 0x5140), t3 = __byte_perm(g23, 0x64646464u, 0x7362);
            mma16816(acc[j], t0, t1, t2, t3, pb_lo, pb_hi);
          }
        } else {
        // ---------------- O^T += (1024 + c n_V)^T . P'^T
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t A = j < 4 ? W0.x : W0.y, B = j < 4 ? W1.x : W1.y, Cw = j < 4 ? W2.x : W2.y,
                         D = j < 4 ? W3.x : W3.y;
          const uint32_t sel = (uint32_t)(j & 3) | ((uint32_t)(4 + (j & 3)) << 8);
          const uint32_t u01 = __byte_perm(A, B, sel), u23 = __byte_perm(Cw, D, sel);
          uint32_t t0, t1, t2, t3;
          asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(t0) : "r"(u01));
          asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(t1) : "r"(u01));
          asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(t2) : "r"(u23));
          asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(t3) : "r"(u23));
          mma16816(acc[j], t0, t1, t2, t3, pb_lo, pb_hi);
        }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == NSTAGE) { s = 0; ph ^= 1; }
    }
    if (tid == 0) OB_AT(4);   // page loop done (warp 0)
    // per-warp totals: l, sp and corr are per-thread partials over the rows g of equal c
#pragma unroll
    for (int k = 4; k <= 16; k <<= 1) {
      l0 += __shfl_xor_sync(0xffffffffu, l0, k);
      l1 += __shfl_xor_sync(0xffffffffu, l1, k);
      corr0 += __shfl_xor_sync(0xffffffffu, corr0, k);
      corr1 += __shfl_xor_sync(0xffffffffu, corr1, k);
      sp0 += __shfl_xor_sync(0xffffffffu, sp0, k);
      sp1 += __shfl_xor_sync(0xffffffffu, sp1, k);
    }
  }
  __syncthreads();  // every stage consumed: the ring can be reused as the merge buffer
  float* obuf = reinterpret_cast<float*>(ring);  // [4 warps][8 heads][128 dims]
  if (warp < 4) {
    const int g = lane >> 2, c = lane & 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // acc[j] = O^T[dims 16g+2j (rows g), 16g+2j+1 (rows g+8)][heads 2c, 2c+1]
      float* o0 = obuf + ((warp * 8 + 2 * c) * DH) + 16 * g + 2 * j;
      float* o1 = o0 + DH;
      if (KV8) {
        const float v_ts = p.kv_scale_quant_orig[1];      // per-tensor dequant scale of V
        o0[0] = (acc[j][0] - 1152.f * sp0) * v_ts; o0[1] = (acc[j][2] - 1152.f * sp0) * v_ts;
        o1[0] = (acc[j][1] - 1152.f * sp1) * v_ts; o1[1] = (acc[j][3] - 1152.f * sp1) * v_ts;
      } else {
        o0[0] = (acc[j][0] - 1024.f * sp0) - corr0; o0[1] = (acc[j][2] - 1024.f * sp0) * 0.0625f - corr0;
        o1[0] = (acc[j][1] - 1024.f * sp1) - corr1; o1[1] = (acc[j][3] - 1024.f * sp1) * 0.0625f - corr1;
      }
    }
    if (g == 0) {
      ml_s[warp][2 * c][0] = m0; ml_s[warp][2 * c][1] = l0;
      ml_s[warp][2 * c + 1][0] = m1; ml_s[warp][2 * c + 1][1] = l1;
    }
  }
  __syncthreads();
  // ------------------------------------------------------------------ merge warps (+ new token), finish / split merge
  const size_t slot = ((size_t)b * gridDim.y + blockIdx.y) * p.n_split + split;
  for (int item = tid; item < G * DH; item += V2_THREADS) {
    const int h = item / DH, d = item - h * DH;
    float M = fmaxf(fmaxf(ml_s[0][h][0], ml_s[1][h][0]), fmaxf(ml_s[2][h][0], ml_s[3][h][0]));
    if (owns_current) M = fmaxf(M, cur_logit_s[h]);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float f = ex2(ml_s[w][h][0] - M);
      num += f * obuf[(w * 8 + h) * DH + d];
      den += f * ml_s[w][h][1];
    }
    if (owns_current) {
      const float f = ex2(cur_logit_s[h] - M);
      num += f * __half2float(kv_new[1][d]);
      den += f;
    }
    if (p.n_split == 1) {
      p.out[((size_t)b * p.Hq + hq0 + h) * DH + d] = __float2half_rn(num * __fdividef(1.f, den + 1.e-6f));
    } else {
      p.part_o[(slot * 8 + h) * DH + d] = num;
      if (d == 0) { p.part_ml[(slot * 8 + h) * 2] = M; p.part_ml[(slot * 8 + h) * 2 + 1] = den; }
    }
  }
  if (p.n_split == 1) {
    if (tid == 0) OB_AT(5);   // outputs of this CTA stored
    if (p.q_out) fused_quant_tail(p, b, &flag_s, red_s);
    if (tid == 0) OB_AT(6);   // exit (the last CTA of a sequence: after the fused row quantisation)
    return;
  }
  __threadfence();
  __syncthreads();
  const int cidx = b * gridDim.y + blockIdx.y;
  if (tid == 0) flag_s = (atomicAdd(&p.counters[cidx], 1) == p.n_split - 1);
  __syncthreads();
  if (!flag_s) return;
  __threadfence();
  const size_t slot0 = ((size_t)b * gridDim.y + blockIdx.y) * p.n_split;
  for (int item = tid; item < G * DH; item += V2_THREADS) {
    const int h = item / DH, d = item - h * DH;
    float gm = -1.0e30f;
    for (int s = 0; s < p.n_split; ++s) gm = fmaxf(gm, __ldcg(&p.part_ml[((slot0 + s) * 8 + h) * 2]));
    float num = 0.f, den = 0.f;
    for (int s = 0; s < p.n_split; ++s) {
      const float w = ex2(__ldcg(&p.part_ml[((slot0 + s) * 8 + h) * 2]) - gm);
      num += w * __ldcg(&p.part_o[((slot0 + s) * 8 + h) * DH + d]);
      den += w * __ldcg(&p.part_ml[((slot0 + s) * 8 + h) * 2 + 1]);
    }
    p.out[((size_t)b * p.Hq + hq0 + h) * DH + d] = __float2half_rn(num * __fdividef(1.f, den + 1.e-6f));
  }
  if (tid == 0) p.counters[cidx] = 0;
  if (p.q_out) fused_quant_tail(p, b, &flag_s, red_s);
}

// ------------------------------------------------------------------------------------------------
// Prefill KV writer: apply_bias_rope_update_kv_cache (no bias)
// /root/reference/kernels/csrc/fused_attention/fused_attention_fine_grained/fine_grained_common/
//   applyBiasRopeUpdateKVCache.h:99-556  -- in-place NeoX RoPE of q and k inside the packed qkv buffer,
// per-token/per-head 4-bit quantisation of post-RoPE k and of v into the pages.
// One warp per (token, head) item; lane l owns elements {2l, 2l+1, 64+2l, 65+2l} = RoPE pairs
// (2l, 2l+64), (2l+1, 2l+65) and page bytes l and 32+l.
// ------------------------------------------------------------------------------------------------
struct PrefillParams {
  __half* qkv; const int* seq_lens; const int* padding_offset; int max_seq_len;
  const int64_t* r_tab; const int64_t* s_tab; int r_max_pages, s_max_pages;
  const int* retrieval_flags; const int* head_rank;
  int T, Hq, Hkv, r_hpool, s_hpool;
  int sink_tok, local_tok, sink_blk, local_blk;
  int rotary_dim; float rope_base, rope_scale;
  const float* kv_scale_orig_quant;   // non-null: per-tensor INT8 pages (K scale, V scale), 128 B per token row
};

// per-tensor INT8 row of one (token, kv head): lane l holds dims 2l, 2l+1, 64+2l, 65+2l (per_tensor_common/update_kv_cache.cu:27-,
// store_8bits_kv_cache_vec: code = cvt.rni.sat.s8(x * scale_orig_quant))
OB_DEVICE void quant8_store_pairs(const float (&x)[4], uint8_t* row, float sq, int lane) {
  const uint16_t lo = (uint16_t)((uint8_t)f2i8_rni_sat(x[0] * sq) | ((uint32_t)(uint8_t)f2i8_rni_sat(x[1] * sq) << 8));
  const uint16_t hi = (uint16_t)((uint8_t)f2i8_rni_sat(x[2] * sq) | ((uint32_t)(uint8_t)f2i8_rni_sat(x[3] * sq) << 8));
  reinterpret_cast<uint16_t*>(row)[lane] = lo;
  reinterpret_cast<uint16_t*>(row + 64)[lane] = hi;
}

OB_DEVICE void quant_store_pairs(const float (&x)[4], uint8_t* row, __half* scale_ptr, __half* zero_ptr, int lane) {
  float mx = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3]));
  float mn = fminf(fminf(x[0], x[1]), fminf(x[2], x[3]));
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, m));
  }
  const __half hs = __float2half_rn(__fdividef(mx - mn, 15.0f));
  const __half hz = __float2half_rn(__fdividef(-15.0f * mn, mx - mn));
  const float inv = __fdividef(1.0f, __half2float(hs));
  const float z = __half2float(hz);
  uint32_t q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = f2u8_rni_sat(__fmaf_rn(x[i], inv, z)) & 0xFu;
  row[lane] = (uint8_t)(q[0] | (q[1] << 4));
  row[32 + lane] = (uint8_t)(q[2] | (q[3] << 4));
  if (lane == 0) { *scale_ptr = hs; *zero_ptr = hz; }
}

__global__ void __launch_bounds__(256) kv4_prefill_write_kernel(const PrefillParams p) {
  // cos / sin depend on (position, pair index) only: they are evaluated once per token (64 accurate sincosf) and
  // shared by the token's 40 q / k heads through shared memory; powf(base, 2d/rot) once per CTA.
  constexpr int TOK = 4;                       // tokens per CTA iteration (256 threads = 4 x 64 pair indices)
  __shared__ float pw[DH / 2];
  __shared__ float cs_s[TOK][DH / 2], sn_s[TOK][DH / 2];
  __shared__ int b_s[TOK], pos_s[TOK], len_s[TOK];
  pdl_trigger();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int heads_total = p.Hq + 2 * p.Hkv;
  const int row_elems = heads_total * DH;
  const int half_rot = p.rotary_dim >> 1;
  if (tid < DH / 2) pw[tid] = powf(p.rope_base, (float)(2 * tid) / (float)p.rotary_dim);
  pdl_wait();
  __syncthreads();
  for (int t0 = blockIdx.x * TOK; t0 < p.T; t0 += gridDim.x * TOK) {
    {
      const int j = tid >> 6, d = tid & 63, t = t0 + j;
      if (t < p.T) {
        const int g = t + p.padding_offset[t];
        const int b = g / p.max_seq_len;
        const int pos = g - b * p.max_seq_len;
        if (d == 0) { b_s[j] = b; pos_s[j] = pos; len_s[j] = p.seq_lens[b]; }
        if (d < half_rot) sincosf(((float)pos * p.rope_scale) / pw[d], &sn_s[j][d], &cs_s[j][d]);
      }
    }
    __syncthreads();
    const int items = min(TOK, p.T - t0) * heads_total;
    for (int it = warp; it < items; it += 8) {
      const int j = it / heads_total;
      const int hh = it - j * heads_total;
      const int t = t0 + j;
      const int b = b_s[j], pos = pos_s[j];
      __half* src = p.qkv + (size_t)t * row_elems + (size_t)hh * DH;
      const __half2 lo = *reinterpret_cast<const __half2*>(src + 2 * lane);
      const __half2 hi = *reinterpret_cast<const __half2*>(src + 64 + 2 * lane);
      float x[4] = {__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi)};
      const bool is_v = hh >= p.Hq + p.Hkv;
      if (!is_v) {
        // pairs (2l, 2l+64) and (2l+1, 2l+65); rotary_dim == 128 on this path (Dh == rotary_dim)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int d = 2 * lane + e;
          if (d < half_rot) {
            const float sn = sn_s[j][d], cs = cs_s[j][d];
            const float a = x[e], bb = x[2 + e];
            x[e] = __half2float(__float2half_rn(cs * a - sn * bb));
            x[2 + e] = __half2float(__float2half_rn(cs * bb + sn * a));
          }
        }
        *reinterpret_cast<__half2*>(src + 2 * lane) = __floats2half2_rn(x[0], x[1]);
        *reinterpret_cast<__half2*>(src + 64 + 2 * lane) = __floats2half2_rn(x[2], x[3]);
      }
      if (hh < p.Hq) continue;
      const int hkv = is_v ? hh - p.Hq - p.Hkv : hh - p.Hq;
      const bool retr = p.retrieval_flags ? p.retrieval_flags[hkv] != 0 : true;
      const int rank = p.head_rank ? p.head_rank[hkv] : hkv;
      const int L = len_s[j];
      int tabidx = pos >> 6;
      const int64_t* tab;
      int hpool;
      if (retr) {
        tab = p.r_tab + (size_t)b * 2 * p.r_max_pages + (is_v ? p.r_max_pages : 0);
        hpool = p.r_hpool;
      } else {
        if (!(pos < p.sink_tok || pos >= L - p.local_tok)) continue;  // applyBiasRopeUpdateKVCache.h:303-311
        tab = p.s_tab + (size_t)b * 2 * p.s_max_pages + (is_v ? p.s_max_pages : 0);
        hpool = p.s_hpool;
        tabidx = tabidx < p.sink_blk ? tabidx : p.sink_blk + (tabidx - p.sink_blk) % p.local_blk;
      }
      uint8_t* page = reinterpret_cast<uint8_t*>(tab[tabidx]);
      const int slot = pos & 63;
      if (p.kv_scale_orig_quant) {
        quant8_store_pairs(x, page + (size_t)rank * TPB * DH + slot * DH, p.kv_scale_orig_quant[is_v ? 1 : 0], lane);
        continue;
      }
      const int data_bytes = hpool * TPB * (DH / 2);
      __half* sc = reinterpret_cast<__half*>(page + data_bytes) + rank * TPB + slot;
      quant_store_pairs(x, page + (size_t)rank * TPB * (DH / 2) + slot * (DH / 2), sc, sc + hpool * TPB, lane);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Fused prefill pass: apply_bias_rope_update_kv_cache + paged_min_max_pool in ONE kernel.
// The reference runs them back to back (omniserve/modeling/layers/ctx_update_kv.py:104-178, llama_w4a8_unpad.py:309-325):
// the writer rotates q / k in place and writes the KV4 pages, then the pool kernel re-reads the rotated keys of the whole
// chunk from HBM to build the kmax / kmin page statistics.  Here the work item is one 16-token sub-chunk of one sequence:
// the CTA rotates and quantises its tokens exactly like kv4_prefill_write_kernel (same arithmetic -> same page bytes), keeps
// the rotated fp16 keys of the retrieval heads in shared memory and reduces them to the sub-chunk's statistics before moving
// on -- the keys are read once, and the statistics land in the page the nibbles just went to.
// ------------------------------------------------------------------------------------------------
constexpr int FP_SUB = 16;        // tokens per sub-chunk (LServe: 64-token pages, 4 sub-chunks)
constexpr int FP_MAX_HR = 8;      // retrieval kv heads kept in shared memory

__global__ void __launch_bounds__(256) kv4_prefill_write_pool_kernel(const PrefillParams p, const int B, const long long stats_off_bytes,
                                                                      const int eles_per_ind) {
  __shared__ float pw[DH / 2];
  __shared__ float cs_s[FP_SUB][DH / 2], sn_s[FP_SUB][DH / 2];
  __shared__ __align__(16) __half kst[FP_SUB][FP_MAX_HR][DH];   // rotated keys of the retrieval heads
  __shared__ int item_s[4];                                       // b, pos0, n, packed start
  pdl_trigger();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int heads_total = p.Hq + 2 * p.Hkv;
  const int row_elems = heads_total * DH;
  const int half_rot = p.rotary_dim >> 1;
  if (tid < DH / 2) pw[tid] = powf(p.rope_base, (float)(2 * tid) / (float)p.rotary_dim);
  pdl_wait();
  // total number of sub-chunk items
  int total = 0;
  for (int b = 0; b < B; ++b) total += (p.seq_lens[b] + FP_SUB - 1) / FP_SUB;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    __syncthreads();
    if (tid == 0) {
      int rem = item, start = 0, b = 0;
      for (; b < B; ++b) {
        const int L = p.seq_lens[b], c = (L + FP_SUB - 1) / FP_SUB;
        if (rem < c) break;
        rem -= c; start += L;
      }
      const int L = p.seq_lens[b];
      item_s[0] = b; item_s[1] = rem * FP_SUB; item_s[2] = min(FP_SUB, L - rem * FP_SUB); item_s[3] = start + rem * FP_SUB;
    }
    __syncthreads();
    const int b = item_s[0], pos0 = item_s[1], n = item_s[2], t0 = item_s[3];
    const int L = p.seq_lens[b];
    for (int i = tid; i < n * (DH / 2); i += 256) {
      const int j = i >> 6, d = i & 63;
      if (d < half_rot) sincosf(((float)(pos0 + j) * p.rope_scale) / pw[d], &sn_s[j][d], &cs_s[j][d]);
    }
    __syncthreads();
    for (int it = warp; it < n * heads_total; it += 8) {
      const int j = it / heads_total;
      const int hh = it - j * heads_total;
      const int pos = pos0 + j;
      __half* src = p.qkv + (size_t)(t0 + j) * row_elems + (size_t)hh * DH;
      const __half2 lo = *reinterpret_cast<const __half2*>(src + 2 * lane);
      const __half2 hi = *reinterpret_cast<const __half2*>(src + 64 + 2 * lane);
      float x[4] = {__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi)};
      const bool is_v = hh >= p.Hq + p.Hkv;
      if (!is_v) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int d = 2 * lane + e;
          if (d < half_rot) {
            const float sn = sn_s[j][d], cs = cs_s[j][d];
            const float a = x[e], bb = x[2 + e];
            x[e] = __half2float(__float2half_rn(cs * a - sn * bb));
            x[2 + e] = __half2float(__float2half_rn(cs * bb + sn * a));
          }
        }
        *reinterpret_cast<__half2*>(src + 2 * lane) = __floats2half2_rn(x[0], x[1]);
        *reinterpret_cast<__half2*>(src + 64 + 2 * lane) = __floats2half2_rn(x[2], x[3]);
      }
      if (hh < p.Hq) continue;
      const int hkv = is_v ? hh - p.Hq - p.Hkv : hh - p.Hq;
      const bool retr = p.retrieval_flags ? p.retrieval_flags[hkv] != 0 : true;
      const int rank = p.head_rank ? p.head_rank[hkv] : hkv;
      if (!is_v && retr) {
        *reinterpret_cast<__half2*>(&kst[j][rank][2 * lane]) = __floats2half2_rn(x[0], x[1]);
        *reinterpret_cast<__half2*>(&kst[j][rank][64 + 2 * lane]) = __floats2half2_rn(x[2], x[3]);
      }
      int tabidx = pos >> 6;
      const int64_t* tab;
      int hpool;
      if (retr) {
        tab = p.r_tab + (size_t)b * 2 * p.r_max_pages + (is_v ? p.r_max_pages : 0);
        hpool = p.r_hpool;
      } else {
        if (!(pos < p.sink_tok || pos >= L - p.local_tok)) continue;  // applyBiasRopeUpdateKVCache.h:303-311
        tab = p.s_tab + (size_t)b * 2 * p.s_max_pages + (is_v ? p.s_max_pages : 0);
        hpool = p.s_hpool;
        tabidx = tabidx < p.sink_blk ? tabidx : p.sink_blk + (tabidx - p.sink_blk) % p.local_blk;
      }
      uint8_t* page = reinterpret_cast<uint8_t*>(tab[tabidx]);
      const int slot = pos & 63;
      const int data_bytes = hpool * TPB * (DH / 2);
      __half* sc = reinterpret_cast<__half*>(page + data_bytes) + rank * TPB + slot;
      quant_store_pairs(x, page + (size_t)rank * TPB * (DH / 2) + slot * (DH / 2), sc, sc + hpool * TPB, lane);
    }
    __syncthreads();
    // statistics of this sub-chunk: channel-wise max / min over its valid tokens (context_pool_kernel.cu:16-69)
    if (p.r_hpool > 0 && p.r_tab) {
      uint8_t* kpage = reinterpret_cast<uint8_t*>(p.r_tab[(size_t)b * 2 * p.r_max_pages + (pos0 >> 6)]);
      const int sub_idx = (pos0 & 63) / FP_SUB;
      __half* kmax = reinterpret_cast<__half*>(kpage + stats_off_bytes) + (size_t)sub_idx * eles_per_ind;
      __half* kmin = kmax + (size_t)(TPB / FP_SUB) * eles_per_ind;
      for (int col = tid; col < p.r_hpool * (DH / 2); col += 256) {
        const int r = col / (DH / 2), d2 = col - r * (DH / 2);
        __half2 mx = *reinterpret_cast<const __half2*>(&kst[0][r][2 * d2]);
        __half2 mn = mx;
        for (int j = 1; j < n; ++j) {
          const __half2 v = *reinterpret_cast<const __half2*>(&kst[j][r][2 * d2]);
          mx = __hmax2(mx, v);
          mn = __hmin2(mn, v);
        }
        *reinterpret_cast<__half2*>(kmax + r * DH + 2 * d2) = mx;
        *reinterpret_cast<__half2*>(kmin + r * DH + 2 * d2) = mn;
      }
    }
  }
}

__global__ void padding_offsets_kernel(int* out, const int* cu, int max_seq_len) {
  const int b = blockIdx.x;
  const int beg = cu[b], end = cu[b + 1];
  const int off = b * max_seq_len - beg;
  for (int t = threadIdx.x; t < end - beg; t += blockDim.x) out[beg + t] = off;
}

// ---------------------------------------------------------------------------------------------- host
// Split-KV scratch: one fixed-size workspace per (device, stream), carved out of a pool that is allocated once on the
// first call on a device and never freed or moved -- pointers baked into captured CUDA graphs stay valid, and two
// streams running attention concurrently never share partials or arrival counters.
constexpr int ATT_MAX_DEV = 16;
constexpr int ATT_MAX_STREAMS = 8;
constexpr size_t ATT_MAX_SLOTS = 4096;          // (sequence, head group, split) partials; 4 KB + 64 B each
constexpr size_t ATT_CNT_INTS = 65536 + 32768;  // [0, 65536): split arrival counters; [65536, ...): fused-quant counters
struct AttWs { cudaStream_t st; float* part_o; float* part_ml; int* cnt; };
struct AttDev { int sms = 0; AttWs ws[ATT_MAX_STREAMS] = {}; int n = 0; };
static AttDev g_att[ATT_MAX_DEV];
static std::mutex g_att_mu;

static int att_sms(int dev) {
  if (!g_att[dev].sms) cudaDeviceGetAttribute(&g_att[dev].sms, cudaDevAttrMultiProcessorCount, dev);
  return g_att[dev].sms;
}

static int get_att_ws(int dev, cudaStream_t st, AttWs* out) {
  std::lock_guard<std::mutex> lk(g_att_mu);
  AttDev& d = g_att[dev];
  if (!d.ws[0].part_o) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone) return OB_ERR_ARG;   // first use must be outside graph capture (warm-up)
    const size_t o_f = ATT_MAX_SLOTS * 8 * DH, ml_f = ATT_MAX_SLOTS * 8 * 2;
    float* po = nullptr; float* pm = nullptr; int* pc = nullptr;
    if (cudaMalloc(&po, o_f * 4 * ATT_MAX_STREAMS) != cudaSuccess) return OB_ERR_CUDA;
    if (cudaMalloc(&pm, ml_f * 4 * ATT_MAX_STREAMS) != cudaSuccess) return OB_ERR_CUDA;
    if (cudaMalloc(&pc, ATT_CNT_INTS * 4 * ATT_MAX_STREAMS) != cudaSuccess) return OB_ERR_CUDA;
    cudaMemset(pc, 0, ATT_CNT_INTS * 4 * ATT_MAX_STREAMS);
    cudaDeviceSynchronize();
    for (int i = 0; i < ATT_MAX_STREAMS; ++i) {
      d.ws[i].part_o = po + (size_t)i * o_f;
      d.ws[i].part_ml = pm + (size_t)i * ml_f;
      d.ws[i].cnt = pc + (size_t)i * ATT_CNT_INTS;
    }
  }
  for (int i = 0; i < d.n; ++i)
    if (d.ws[i].st == st) { *out = d.ws[i]; return 0; }
  if (d.n == ATT_MAX_STREAMS) return OB_ERR_ARG;
  d.ws[d.n].st = st;
  *out = d.ws[d.n++];
  return 0;
}

int kv4_decode_run(const KV4DecodeArgs& a, cudaStream_t st) {
  if (a.B <= 0) return 0;
  if (a.head_dim != DH || a.tokens_per_block != TPB) return OB_ERR_SHAPE;
  if (a.Hq % a.Hkv) return OB_ERR_SHAPE;
  const int group = a.Hq / a.Hkv;
  const bool per_q = a.dyn_idx != nullptr;
  const int G = per_q ? 1 : group;
  if (G < 1 || G > 8) return OB_ERR_SHAPE;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= ATT_MAX_DEV) return OB_ERR_ARG;
  const int g_att_sms = att_sms(dev);
  AttWs ws{};

  AttnParams p{};
  p.q = a.q; p.k = a.k; p.v = a.v; p.q_bs = a.q_bs; p.k_bs = a.k_bs; p.v_bs = a.v_bs; p.out = a.out;
  p.r_tab = a.retrieval_kv_pointers; p.s_tab = a.streaming_kv_pointers;
  p.r_max_pages = a.r_max_pages; p.s_max_pages = a.s_max_pages;
  p.lengths = a.lengths; p.retrieval_flags = a.retrieval_head_flags; p.head_rank = a.head_rank_table;
  p.dyn_idx = a.dyn_idx; p.dyn_pages = a.dyn_pages;
  p.B = a.B; p.Hq = a.Hq; p.Hkv = a.Hkv;
  p.r_hpool = a.num_retrieval_kv_heads; p.s_hpool = a.num_streaming_kv_heads;
  p.sink_tok = a.sink_tokens; p.local_tok = a.local_tokens; p.sink_blk = a.sink_blocks;
  p.local_blk = a.local_blocks > 0 ? a.local_blocks : 1;
  p.rope_base = a.rotary_base; p.rope_scale = a.rotary_scale; p.rotary_dim = a.rotary_dim;
  p.timestep = a.timestep;
  p.sub_chunk = a.tokens_per_sub_chunk; p.eles_per_ind = a.hidden_dim_per_retrieval_token;
  p.q_out = a.q_out; p.q_scale = a.q_scale; p.q_sum = a.q_sum; p.tok_counters = nullptr;
  p.stable_history = a.stable_history;
#ifdef OB_ATT_TIMING
  {
    const char* e = getenv("OB_ATT_DBGT");
    p.dbg_t = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 10)) : nullptr;
    const char* m = getenv("OB_ATT_DBG");      // read at every launch: the tool switches modes between graphs
    p.dbg_mode = m ? atoi(m) : 0;
  }
#endif
  if (p.q_out) {
    if (!p.q_scale || (a.Hq * DH) % 8 || (a.Hq * DH) / 8 > 8 * 128 || a.B > 32768) return OB_ERR_SHAPE;
    if (int e = get_att_ws(dev, st, &ws)) return e;
    p.tok_counters = ws.cnt + 65536;
  }
  if (p.sub_chunk < 0 || (p.sub_chunk > 0 && TPB % p.sub_chunk)) return OB_ERR_SHAPE;

  // KV splits: enough CTAs to balance 148 SMs x ~4 resident CTAs, at least 4 pages per split
  const int max_pages = std::max(1, (std::max(1, a.max_attended) + TPB - 1) / TPB);
  const int ctas_y = per_q ? a.Hq : a.Hkv;
  const int base_ctas = a.B * ctas_y;
  int n_split = 1;
  const int slots = 4 * g_att_sms;  // 4 CTAs of 160 threads are resident per SM
  if (base_ctas * 5 < slots * 4)   // otherwise one CTA per (sequence, kv head) already fills >= 80 % of the machine
    while (base_ctas * n_split < slots && max_pages / (n_split + 1) >= 4 && n_split < 64) ++n_split;
  if (a.force_split > 0) n_split = a.force_split;
  p.n_split = n_split;
  if (n_split > 1) {
    // the workspace is fixed-size: automatic splits stay below ~2 * 4 * #SMs partials; a forced split that needs more
    // is rejected rather than growing (and moving) the buffer
    if (base_ctas > 65536 || (size_t)base_ctas * n_split > ATT_MAX_SLOTS) return OB_ERR_ARG;
    if (!ws.part_o) { if (int e = get_att_ws(dev, st, &ws)) return e; }
    p.part_o = ws.part_o; p.part_ml = ws.part_ml; p.counters = ws.cnt;
  }
  dim3 grid(n_split, ctas_y, a.B);
  if (a.kv_scale_quant_orig) {
    if (!a.kv_scale_orig_quant) return OB_ERR_ARG;
    p.kv_scale_quant_orig = a.kv_scale_quant_orig; p.kv_scale_orig_quant = a.kv_scale_orig_quant;
    static bool attr_done[ATT_MAX_DEV] = {};
    if (!attr_done[dev]) {
      if (cudaFuncSetAttribute(kv4_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, V8_STAGES * V8_STAGE_BYTES) != cudaSuccess)
        return OB_ERR_CUDA;
      attr_done[dev] = true;
    }
    return launch_pdl(kv4_decode_kernel<true>, grid, dim3(V2_THREADS), (size_t)(V8_STAGES * V8_STAGE_BYTES), st, p, G) == cudaSuccess
               ? 0 : OB_ERR_CUDA;
  }
  return launch_pdl(kv4_decode_kernel<false>, grid, dim3(V2_THREADS), (size_t)(V2_STAGES * V2_STAGE_BYTES), st, p, G) == cudaSuccess
             ? 0 : OB_ERR_CUDA;
}

}  // namespace ob

namespace ob {

int kv4_prefill_write_run(const KV4PrefillArgs& a, cudaStream_t st) {
  if (a.T <= 0) return 0;
  if (a.rotary_dim != DH) return OB_ERR_SHAPE;
  PrefillParams p{};
  p.qkv = a.qkv; p.seq_lens = a.seq_lens; p.padding_offset = a.padding_offset; p.max_seq_len = a.max_seq_len;
  p.r_tab = a.retrieval_kv_pointers; p.s_tab = a.streaming_kv_pointers;
  p.r_max_pages = a.r_max_pages; p.s_max_pages = a.s_max_pages;
  p.retrieval_flags = a.retrieval_head_flags; p.head_rank = a.head_rank_table;
  p.T = a.T; p.Hq = a.Hq; p.Hkv = a.Hkv; p.r_hpool = a.num_retrieval_kv_heads; p.s_hpool = a.num_streaming_kv_heads;
  p.sink_tok = a.sink_tokens; p.local_tok = a.local_tokens; p.sink_blk = a.sink_blocks;
  p.local_blk = a.local_blocks > 0 ? a.local_blocks : 1;
  p.rotary_dim = a.rotary_dim; p.rope_base = a.rotary_base; p.rope_scale = a.rotary_scale;
  p.kv_scale_orig_quant = a.kv_scale_orig_quant;
  const int blocks = (int)std::min<long long>(((long long)a.T + 3) / 4, 148LL * 8);
  return launch_pdl(kv4_prefill_write_kernel, dim3(blocks), dim3(256), 0, st, p) == cudaSuccess ? 0 : OB_ERR_CUDA;
}

int padding_offsets_run(int* out, const int* cu_seqlens, int B, int max_seq_len, cudaStream_t st) {
  if (B <= 0) return 0;
  padding_offsets_kernel<<<B, 256, 0, st>>>(out, cu_seqlens, max_seq_len);
  return cudaGetLastError() == cudaSuccess ? 0 : OB_ERR_CUDA;
}

int kv4_prefill_write_pool_run(const KV4PrefillArgs& a, int tokens_per_sub_chunk, cudaStream_t st) {
  if (a.T <= 0) return 0;
  if (a.rotary_dim != DH || tokens_per_sub_chunk != FP_SUB || a.num_retrieval_kv_heads > FP_MAX_HR || a.B <= 0) return OB_ERR_SHAPE;
  if (a.kv_scale_orig_quant) return OB_ERR_ARG;   // the fused statistics pass is KV4-only
  PrefillParams p{};
  p.qkv = a.qkv; p.seq_lens = a.seq_lens; p.padding_offset = a.padding_offset; p.max_seq_len = a.max_seq_len;
  p.r_tab = a.retrieval_kv_pointers; p.s_tab = a.streaming_kv_pointers;
  p.r_max_pages = a.r_max_pages; p.s_max_pages = a.s_max_pages;
  p.retrieval_flags = a.retrieval_head_flags; p.head_rank = a.head_rank_table;
  p.T = a.T; p.Hq = a.Hq; p.Hkv = a.Hkv; p.r_hpool = a.num_retrieval_kv_heads; p.s_hpool = a.num_streaming_kv_heads;
  p.sink_tok = a.sink_tokens; p.local_tok = a.local_tokens; p.sink_blk = a.sink_blocks;
  p.local_blk = a.local_blocks > 0 ? a.local_blocks : 1;
  p.rotary_dim = a.rotary_dim; p.rope_base = a.rotary_base; p.rope_scale = a.rotary_scale;
  // K page: [H][64][64] nibbles | scales f16 [H][64] | zeros f16 [H][64] | kmax f16 [4][H*128] | kmin
  const long long stats_off = (long long)a.num_retrieval_kv_heads * TPB * (DH / 2) + (long long)a.num_retrieval_kv_heads * TPB * 4;
  const int eles = a.num_retrieval_kv_heads * DH;
  const int items = (a.T + FP_SUB - 1) / FP_SUB + a.B;
  const int blocks = std::min(items, 148 * 4);
  return launch_pdl(kv4_prefill_write_pool_kernel, dim3(blocks), dim3(256), 0, st, p, a.B, stats_off, eles) == cudaSuccess
             ? 0 : OB_ERR_CUDA;
}

}  // namespace ob
