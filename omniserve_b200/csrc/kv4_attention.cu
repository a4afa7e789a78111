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
//   * one CTA per (sequence, KV head, KV split) serves the whole GQA group, so K/V nibbles are read
//     from HBM once per KV head instead of once per query head;
//   * 16-byte coalesced loads (4 lanes x 16 B per token row for K, 8 lanes x 8 B for V);
//   * the per-token scale/zero are folded out of the inner loops:
//       q.k = s*(sum_d q_d (n_d-8)) + s*(8-z)*sum_d q_d ;  sum_t p_t v_t = sum_t (p_t s_t)(n-8) - sum_t p_t s_t (z_t-8)
//     so the loops work on exact small integers in half2 and never touch a dequant FMA;
//   * P.V accumulates in half2 for 8 tokens and is flushed to fp32 (reference: fp32 throughout);
//   * split-KV with (max,sum,out) partials merged by the last-arriving CTA (flash-decoding), so the
//     context length is not bounded by shared memory.
// Numerics therefore agree with the reference to ~1e-3 (north-star tolerance), not bit-exactly; the KV
// page bytes written for the new token follow the reference formula exactly.
#include "kv4_attention.h"
#include "ptx.cuh"

#include <algorithm>
#include <float.h>

namespace ob {

constexpr int DH = 128;
constexpr int TPB = 64;            // tokens per page
constexpr int ATT_THREADS = 128;
constexpr int MAX_CHUNK = 4096;    // cached tokens handled by one CTA (logits live in smem)

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
  OB_DEVICE int pos_of(int i) const {
    if (mode == 1) return i < sink_tok ? i : i + gap;
    if (mode == 2) return dyn[i >> 6] * TPB + (i & 63);
    return i;
  }
  OB_DEVICE int tab_idx(int pos) const {
    int blk = pos >> 6;
    if (mode == 1) blk = blk < sink_blk ? blk : sink_blk + (blk - sink_blk) % local_blk;
    return blk;
  }
};

// (n-8) as half2 pairs (n_j, n_{j+4}), j = 0..3, of one 32-bit word of 8 nibbles.
OB_DEVICE void nib8_to_h2(uint32_t w, __half2 (&o)[4]) {
  uint32_t t0, t1, t2, t3;
  const uint32_t top = w >> 8;
  asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(t0) : "r"(w));
  asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(t1) : "r"(w));
  asm("lop3.b32 %0, %1, 0x000f000f, 0x64006400, 0xea;" : "=r"(t2) : "r"(top));
  asm("lop3.b32 %0, %1, 0x00f000f0, 0x64006400, 0xea;" : "=r"(t3) : "r"(top));
  const __half2 c1032 = __halves2half2(__ushort_as_half(0x6408), __ushort_as_half(0x6408));
  const __half2 c16th = __halves2half2(__ushort_as_half(0x2c00), __ushort_as_half(0x2c00));
  const __half2 cm72 = __halves2half2(__ushort_as_half(0xd480), __ushort_as_half(0xd480));
  o[0] = __hsub2(*reinterpret_cast<__half2*>(&t0), c1032);
  o[1] = __hfma2(*reinterpret_cast<__half2*>(&t1), c16th, cm72);
  o[2] = __hsub2(*reinterpret_cast<__half2*>(&t2), c1032);
  o[3] = __hfma2(*reinterpret_cast<__half2*>(&t3), c16th, cm72);
}

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
};

template <int G>
__global__ void __launch_bounds__(ATT_THREADS)
kv4_decode_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // layout: logits float [G][chunk_cap+1] | q_s half [G][128] | red float [...]
  __shared__ __align__(16) __half q_s[G][DH];
  __shared__ __align__(16) __half kv_new[2][DH];
  __shared__ float red[4][G][2];
  __shared__ float qsum_s[G], cur_logit_s[G], stat_s[G][2];
  __shared__ float o_red[4][G][DH];
  __shared__ int flag_s;
  float* logits = reinterpret_cast<float*>(smem_raw);

  const int split = blockIdx.x;
  const int b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int group = p.Hq / p.Hkv;
  // G == group: one CTA per kv head; G == 1: one CTA per q head (dynamic page lists differ per q head)
  const int hq0 = (G == 1) ? blockIdx.y : blockIdx.y * G;
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
      sv.n_valid = (p.dyn_pages - 1) * TPB + (tl - 1) % TPB + 1;
      if (tl <= 0) sv.n_valid = 0;
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
  sv.data_bytes = sv.hpool * TPB * (DH / 2);

  // this CTA's slice of the attended (logical) token list
  const int per_split = ((sv.n_valid + p.n_split - 1) / p.n_split + 31) & ~31;
  const int i0 = min(split * per_split, sv.n_valid);
  const int i1 = min(i0 + per_split, sv.n_valid);
  const int n_loc = i1 - i0;
  const bool owns_current = (split == p.n_split - 1);
  const int cap = per_split + 1;  // logits row pitch

  // ------------------------------------------------------------------ prologue: q/k RoPE, append new K/V
  {
    const float pos = (float)tl;
    // rotate q heads and k: thread d < 64 handles the NeoX pair (d, d + rot/2)
    const int half_rot = p.rotary_dim >> 1;
    for (int item = tid; item < (G + 1) * (DH / 2); item += ATT_THREADS) {
      const int h = item / (DH / 2), d = item - h * (DH / 2);
      const __half* src = (h < G) ? p.q + (size_t)b * p.q_bs + (size_t)(hq0 + h) * DH
                                  : p.k + (size_t)b * p.k_bs + (size_t)hkv * DH;
      __half* dst = (h < G) ? q_s[h] : kv_new[0];
      // generic mapping of the 64 "pair slots": rotary pairs first, then pass-through elements
      if (d < half_rot) {
        const float inv_freq = (pos * p.rope_scale) / powf(p.rope_base, (float)(2 * d) / (float)p.rotary_dim);
        float sn, cs;
        sincosf(inv_freq, &sn, &cs);
        const float x = __half2float(src[d]), y = __half2float(src[d + half_rot]);
        dst[d] = __float2half_rn(cs * x - sn * y);
        dst[d + half_rot] = __float2half_rn(cs * y + sn * x);
      } else {
        const int e = p.rotary_dim + 2 * (d - half_rot);
        dst[e] = src[e];
        dst[e + 1] = src[e + 1];
      }
    }
    for (int d = tid; d < DH; d += ATT_THREADS) kv_new[1][d] = p.v[(size_t)b * p.v_bs + (size_t)hkv * DH + d];
  }
  __syncthreads();
  for (int h = warp; h < G; h += ATT_THREADS / 32) {
    // qsum and the full-precision logit of the new token (Template.hpp:1356-1376)
    float s = 0.f, dot = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float qv = __half2float(q_s[h][lane * 4 + i]);
      s += qv;
      dot += qv * __half2float(kv_new[0][lane * 4 + i]);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, m);
      dot += __shfl_xor_sync(0xffffffffu, dot, m);
    }
    if (lane == 0) { qsum_s[h] = s; cur_logit_s[h] = dot * 0.08838834764831845f; }
  }
  // append (only once per kv head: the split that owns the current token, first q head of the group)
  const bool writer = owns_current && ((G > 1) || (hq0 == hkv * group));
  if (writer && warp >= 2) {
    const int which = warp - 2;  // 0 = K, 1 = V
    const int64_t* tab = which ? sv.vtab : sv.ktab;
    // streaming heads write through the ring mapping, retrieval heads at the true page
    SeqView wv = sv;
    if (wv.mode == 2) wv.mode = 0;
    uint8_t* page = reinterpret_cast<uint8_t*>(tab[wv.tab_idx(tl)]);
    const int slot = tl & 63;
    __half x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = kv_new[which][lane * 4 + i];
    __half* sc = reinterpret_cast<__half*>(page + sv.data_bytes) + sv.rank * TPB + slot;
    quant_store_token(x, page + (size_t)sv.rank * TPB * (DH / 2) + slot * (DH / 2), sc, sc + sv.hpool * TPB, lane);
  }
  __syncthreads();

  // ------------------------------------------------------------------ pass 1: logits = q.K^T
  const float inv_sqrt = 0.08838834764831845f;  // 1/sqrt(128)
  float lmax[G];
#pragma unroll
  for (int h = 0; h < G; ++h) lmax[h] = -FLT_MAX;
  {
    const int c = tid & 3, ts = tid >> 2;
    // q fragments in the nibble-pair order: pairs (8w+j, 8w+j+4) of this lane's 32 dims
    __half2 qf[G][16];
#pragma unroll
    for (int h = 0; h < G; ++h)
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          qf[h][w * 4 + j] = __halves2half2(q_s[h][c * 32 + w * 8 + j], q_s[h][c * 32 + w * 8 + j + 4]);
    float qs[G];
#pragma unroll
    for (int h = 0; h < G; ++h) qs[h] = qsum_s[h];

    for (int base = 0; base < n_loc; base += 32) {
      const int li = base + ts;
      const bool ok = li < n_loc;
      const int pos = sv.pos_of(i0 + (ok ? li : 0));
      const uint8_t* page = reinterpret_cast<const uint8_t*>(sv.ktab[sv.tab_idx(pos)]);
      const int slot = pos & 63;
      const uint4 kw = ld_nc_v4(page + (size_t)sv.rank * TPB * (DH / 2) + slot * (DH / 2) + c * 16);
      const __half* scp = reinterpret_cast<const __half*>(page + sv.data_bytes) + sv.rank * TPB + slot;
      const float ks = __half2float(scp[0]);
      const float kz = __half2float(scp[sv.hpool * TPB]);
      __half2 n2[16];
      {
        __half2 t[4];
        nib8_to_h2(kw.x, t); n2[0] = t[0]; n2[1] = t[1]; n2[2] = t[2]; n2[3] = t[3];
        nib8_to_h2(kw.y, t); n2[4] = t[0]; n2[5] = t[1]; n2[6] = t[2]; n2[7] = t[3];
        nib8_to_h2(kw.z, t); n2[8] = t[0]; n2[9] = t[1]; n2[10] = t[2]; n2[11] = t[3];
        nib8_to_h2(kw.w, t); n2[12] = t[0]; n2[13] = t[1]; n2[14] = t[2]; n2[15] = t[3];
      }
#pragma unroll
      for (int h = 0; h < G; ++h) {
        __half2 a0 = __hmul2(qf[h][0], n2[0]), a1 = __hmul2(qf[h][8], n2[8]);
#pragma unroll
        for (int i = 1; i < 8; ++i) {
          a0 = __hfma2(qf[h][i], n2[i], a0);
          a1 = __hfma2(qf[h][8 + i], n2[8 + i], a1);
        }
        const float2 f0 = __half22float2(a0), f1 = __half22float2(a1);
        float s = (f0.x + f0.y) + (f1.x + f1.y);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const float lg = (ks * s + ks * (8.0f - kz) * qs[h]) * inv_sqrt;
        if (ok && c == 0) {
          logits[h * cap + li] = lg;
          lmax[h] = fmaxf(lmax[h], lg);
        }
      }
    }
  }
  // ------------------------------------------------------------------ softmax statistics
  float lsum[G];
#pragma unroll
  for (int h = 0; h < G; ++h) {
    float m = lmax[h];
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
    if (lane == 0) red[warp][h][0] = m;
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < G; ++h) {
    float m = fmaxf(fmaxf(red[0][h][0], red[1][h][0]), fmaxf(red[2][h][0], red[3][h][0]));
    if (owns_current) m = fmaxf(m, cur_logit_s[h]);
    lmax[h] = m;
    float s = 0.f;
    for (int i = tid; i < n_loc; i += ATT_THREADS) {
      const float e = __expf(logits[h * cap + i] - m);
      logits[h * cap + i] = e;
      s += e;
    }
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if (lane == 0) red[warp][h][1] = s;
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < G; ++h) {
    float s = red[0][h][1] + red[1][h][1] + red[2][h][1] + red[3][h][1];
    if (owns_current) s += __expf(cur_logit_s[h] - lmax[h]);
    lsum[h] = s;
  }
  // n_split == 1: probabilities are normalised and rounded to fp16 before P.V like the reference (:1819-1831)
  float pscale[G];
#pragma unroll
  for (int h = 0; h < G; ++h) pscale[h] = (p.n_split == 1) ? __fdividef(1.f, lsum[h] + 1.e-6f) : 1.f;

  // ------------------------------------------------------------------ pass 2: out = P.V
  float of[G][16];
  float corr[G];
#pragma unroll
  for (int h = 0; h < G; ++h) {
    corr[h] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) of[h][i] = 0.f;
  }
  {
    const int c8 = tid & 7, ts = tid >> 3;
    __half2 o2[G][8];
#pragma unroll
    for (int h = 0; h < G; ++h)
#pragma unroll
      for (int i = 0; i < 8; ++i) o2[h][i] = __float2half2_rn(0.f);
    int since_flush = 0;
    for (int base = 0; base < n_loc; base += 16) {
      const int li = base + ts;
      const bool ok = li < n_loc;
      const int pos = sv.pos_of(i0 + (ok ? li : 0));
      const uint8_t* page = reinterpret_cast<const uint8_t*>(sv.vtab[sv.tab_idx(pos)]);
      const int slot = pos & 63;
      const uint2 vw = ld_nc_v2(page + (size_t)sv.rank * TPB * (DH / 2) + slot * (DH / 2) + c8 * 8);
      const __half* scp = reinterpret_cast<const __half*>(page + sv.data_bytes) + sv.rank * TPB + slot;
      const float vs = ok ? __half2float(scp[0]) : 0.f;
      const float vz = __half2float(scp[sv.hpool * TPB]);
      __half2 n2[8];
      {
        __half2 t[4];
        nib8_to_h2(vw.x, t); n2[0] = t[0]; n2[1] = t[1]; n2[2] = t[2]; n2[3] = t[3];
        nib8_to_h2(vw.y, t); n2[4] = t[0]; n2[5] = t[1]; n2[6] = t[2]; n2[7] = t[3];
      }
#pragma unroll
      for (int h = 0; h < G; ++h) {
        float pr = ok ? logits[h * cap + li] * pscale[h] : 0.f;
        pr = __half2float(__float2half_rn(pr));        // probabilities are fp16 in the reference
        const float ps = pr * vs;
        corr[h] += ps * (vz - 8.0f);
        const __half2 p2 = __float2half2_rn(ps);
#pragma unroll
        for (int i = 0; i < 8; ++i) o2[h][i] = __hfma2(p2, n2[i], o2[h][i]);
      }
      if (++since_flush == 8) {
        since_flush = 0;
#pragma unroll
        for (int h = 0; h < G; ++h)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 f = __half22float2(o2[h][i]);
            of[h][2 * i] += f.x;
            of[h][2 * i + 1] += f.y;
            o2[h][i] = __float2half2_rn(0.f);
          }
      }
    }
#pragma unroll
    for (int h = 0; h < G; ++h)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 f = __half22float2(o2[h][i]);
        of[h][2 * i] += f.x;
        of[h][2 * i + 1] += f.y;
      }
    // reduce over the 4 token slots of this warp (lanes with equal c8), then over warps
#pragma unroll
    for (int h = 0; h < G; ++h) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float x = of[h][i];
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        of[h][i] = x;
      }
      float cc = corr[h];
      // corr is identical on the 8 lanes of a token slot; sum the 4 slots
      cc += __shfl_xor_sync(0xffffffffu, cc, 8);
      cc += __shfl_xor_sync(0xffffffffu, cc, 16);
      corr[h] = cc;
    }
    if (lane < 8) {
      // of[h][2*i + e] holds dim  c8*16 + (i/4)*8 + (i%4) + 4*e
#pragma unroll
      for (int h = 0; h < G; ++h)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int d = c8 * 16 + (i >> 2) * 8 + (i & 3);
          o_red[warp][h][d] = of[h][2 * i] - corr[h];
          o_red[warp][h][d + 4] = of[h][2 * i + 1] - corr[h];
        }
    }
  }
  __syncthreads();
  // ------------------------------------------------------------------ finish / split merge
  for (int item = tid; item < G * DH; item += ATT_THREADS) {
    const int h = item / DH, d = item - h * DH;
    float o = o_red[0][h][d] + o_red[1][h][d] + o_red[2][h][d] + o_red[3][h][d];
    o_red[0][h][d] = o;
  }
  if (tid < G) { stat_s[tid][0] = lmax[tid]; stat_s[tid][1] = lsum[tid]; }
  __syncthreads();
  if (p.n_split == 1) {
    for (int item = tid; item < G * DH; item += ATT_THREADS) {
      const int h = item / DH, d = item - h * DH;
      const float inv = __fdividef(1.f, stat_s[h][1] + 1.e-6f);
      float pc = __expf(cur_logit_s[h] - stat_s[h][0]) * inv;
      pc = __half2float(__float2half_rn(pc));
      const float o = o_red[0][h][d] + pc * __half2float(kv_new[1][d]);
      p.out[((size_t)b * p.Hq + hq0 + h) * DH + d] = __float2half_rn(o);
    }
    return;
  }
  // partials: unnormalised out (relative to local max), local max, local sum
  {
    const size_t slot = ((size_t)b * gridDim.y + blockIdx.y) * p.n_split + split;
    for (int item = tid; item < G * DH; item += ATT_THREADS) {
      const int h = item / DH, d = item - h * DH;
      float o = o_red[0][h][d];
      if (owns_current) o += __expf(cur_logit_s[h] - stat_s[h][0]) * __half2float(kv_new[1][d]);
      p.part_o[(slot * G + h) * DH + d] = o;
    }
    if (tid < G) {
      p.part_ml[(slot * G + tid) * 2] = stat_s[tid][0];
      p.part_ml[(slot * G + tid) * 2 + 1] = stat_s[tid][1];
    }
    __threadfence();
    __syncthreads();
    const int cidx = b * gridDim.y + blockIdx.y;
    if (tid == 0) flag_s = (atomicAdd(&p.counters[cidx], 1) == p.n_split - 1);
    __syncthreads();
    if (!flag_s) return;
    __threadfence();
    const size_t slot0 = ((size_t)b * gridDim.y + blockIdx.y) * p.n_split;
    for (int item = tid; item < G * DH; item += ATT_THREADS) {
      const int h = item / DH, d = item - h * DH;
      float gm = -FLT_MAX;
      for (int s = 0; s < p.n_split; ++s) gm = fmaxf(gm, __ldcg(&p.part_ml[((slot0 + s) * G + h) * 2]));
      float num = 0.f, den = 0.f;
      for (int s = 0; s < p.n_split; ++s) {
        const float w = __expf(__ldcg(&p.part_ml[((slot0 + s) * G + h) * 2]) - gm);
        num += w * __ldcg(&p.part_o[((slot0 + s) * G + h) * DH + d]);
        den += w * __ldcg(&p.part_ml[((slot0 + s) * G + h) * 2 + 1]);
      }
      p.out[((size_t)b * p.Hq + hq0 + h) * DH + d] = __float2half_rn(num * __fdividef(1.f, den + 1.e-6f));
    }
    if (tid == 0) p.counters[cidx] = 0;
  }
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
};

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
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const int heads_total = p.Hq + 2 * p.Hkv;
  const long long items = (long long)p.T * heads_total;
  const int row_elems = heads_total * DH;
  const int half_rot = p.rotary_dim >> 1;
  for (long long it = (long long)blockIdx.x * warps_per_cta + (threadIdx.x >> 5); it < items;
       it += (long long)gridDim.x * warps_per_cta) {
    const int t = (int)(it / heads_total);
    const int hh = (int)(it - (long long)t * heads_total);
    const int g = t + p.padding_offset[t];
    const int b = g / p.max_seq_len;
    const int pos = g - b * p.max_seq_len;
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
          const float inv_freq = ((float)pos * p.rope_scale) / powf(p.rope_base, (float)(2 * d) / (float)p.rotary_dim);
          float sn, cs;
          sincosf(inv_freq, &sn, &cs);
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
    const int L = p.seq_lens[b];
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
}

__global__ void padding_offsets_kernel(int* out, const int* cu, int max_seq_len) {
  const int b = blockIdx.x;
  const int beg = cu[b], end = cu[b + 1];
  const int off = b * max_seq_len - beg;
  for (int t = threadIdx.x; t < end - beg; t += blockDim.x) out[beg + t] = off;
}

// ---------------------------------------------------------------------------------------------- host
static float* g_part_o[16] = {nullptr};
static float* g_part_ml[16] = {nullptr};
static int* g_att_cnt[16] = {nullptr};
static size_t g_part_cap[16] = {0};
static int g_att_sms = 0;

static int ensure_att_ws(int dev, size_t slots, int G) {
  const size_t need = slots * G;
  if (g_part_cap[dev] >= need && g_att_cnt[dev]) return 0;
  if (g_part_o[dev]) { cudaFree(g_part_o[dev]); cudaFree(g_part_ml[dev]); }
  if (cudaMalloc(&g_part_o[dev], need * DH * 4) != cudaSuccess) return OB_ERR_CUDA;
  if (cudaMalloc(&g_part_ml[dev], need * 2 * 4) != cudaSuccess) return OB_ERR_CUDA;
  g_part_cap[dev] = need;
  if (!g_att_cnt[dev]) {
    if (cudaMalloc(&g_att_cnt[dev], 65536 * 4) != cudaSuccess) return OB_ERR_CUDA;
    cudaMemset(g_att_cnt[dev], 0, 65536 * 4);
    cudaDeviceSynchronize();
  }
  return 0;
}

int kv4_decode_run(const KV4DecodeArgs& a, cudaStream_t st) {
  if (a.B <= 0) return 0;
  if (a.head_dim != DH || a.tokens_per_block != TPB) return OB_ERR_SHAPE;
  if (a.Hq % a.Hkv) return OB_ERR_SHAPE;
  const int group = a.Hq / a.Hkv;
  const bool per_q = a.dyn_idx != nullptr;
  const int G = per_q ? 1 : group;
  if (G != 1 && G != 2 && G != 4 && G != 8) return OB_ERR_SHAPE;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!g_att_sms) cudaDeviceGetAttribute(&g_att_sms, cudaDevAttrMultiProcessorCount, dev);

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

  // upper bound of attended cached tokens (host knows only the max): timestep
  const int max_ctx = std::max(1, a.max_attended);
  const int ctas_y = per_q ? a.Hq : a.Hkv;
  int n_split = (max_ctx + MAX_CHUNK - 1) / MAX_CHUNK;
  // fill the machine: aim for >= 2 CTAs per SM when the batch is small
  const int base_ctas = a.B * ctas_y;
  while (base_ctas * n_split < 2 * g_att_sms && max_ctx / (n_split + 1) >= 256) ++n_split;
  if (a.force_split > 0) n_split = a.force_split;
  p.n_split = n_split;
  const int per_split = (((max_ctx + n_split - 1) / n_split) + 31) & ~31;
  const size_t smem = (size_t)G * (per_split + 1) * 4;
  if (smem > 160 * 1024) return OB_ERR_SHAPE;
  if (n_split > 1) {
    if (base_ctas > 65536) return OB_ERR_SHAPE;
    if (int e = ensure_att_ws(dev, (size_t)base_ctas * n_split, G)) return e;
    p.part_o = g_part_o[dev]; p.part_ml = g_part_ml[dev]; p.counters = g_att_cnt[dev];
  }
  dim3 grid(n_split, ctas_y, a.B);
#define OB_ATT(g)                                                                                         \
  case g: {                                                                                               \
    static size_t set = 0;                                                                                \
    if (smem > 48 * 1024 && smem > set) {                                                                 \
      if (cudaFuncSetAttribute(kv4_decode_kernel<g>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                               (int)smem) != cudaSuccess)                                                 \
        return OB_ERR_CUDA;                                                                               \
      set = smem;                                                                                         \
    }                                                                                                     \
    kv4_decode_kernel<g><<<grid, ATT_THREADS, smem, st>>>(p);                                             \
    break;                                                                                                \
  }
  switch (G) {
    OB_ATT(1)
    OB_ATT(2)
    OB_ATT(4)
    OB_ATT(8)
  }
#undef OB_ATT
  return cudaGetLastError() == cudaSuccess ? 0 : OB_ERR_CUDA;
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
  const long long items = (long long)a.T * (a.Hq + 2 * a.Hkv);
  const int blocks = (int)std::min<long long>((items + 7) / 8, 148LL * 16);
  kv4_prefill_write_kernel<<<blocks, 256, 0, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : OB_ERR_CUDA;
}

int padding_offsets_run(int* out, const int* cu_seqlens, int B, int max_seq_len, cudaStream_t st) {
  if (B <= 0) return 0;
  padding_offsets_kernel<<<B, 256, 0, st>>>(out, cu_seqlens, max_seq_len);
  return cudaGetLastError() == cudaSuccess ? 0 : OB_ERR_CUDA;
}

}  // namespace ob
