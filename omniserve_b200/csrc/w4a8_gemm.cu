// W4A8 GEMM for sm_100a: packed-INT4 weights -> INT8 in registers -> tensor memory (A operand) ->
// tcgen05.mma kind::i8 with INT32 accumulators in TMEM; activations TMA-staged to 128B-swizzled shared
// memory (B operand); QServe dequant fused in the epilogue.
//
// Replaces (same math, new design):
//   per-channel: /root/reference/kernels/csrc/qgemm/w4a8_per_chn/gemm_cuda.cu:308-657
//   per-group  : /root/reference/kernels/csrc/qgemm/w4a8_per_group/gemm_cuda.cu:333-707
//
// Orientation: C^T = W . X^T.  The 128 rows of a weight tile are the UMMA M dimension (TMEM lanes),
// BN tokens are the UMMA N dimension (TMEM columns).  The reference's packed weight layout
// [N/32][K/32][32 lanes][16 B] (w4a8_linear.py:297-327) is consumed *unchanged*: the 16 bytes a
// "lane" owns are exactly the four registers of a `tcgen05.st.16x128b.x2` fragment
// (rows c / c+8, K-bytes 4e.. / 16+4e..), low nibbles = rows n, high nibbles = rows n+16.
//
// Warp roles (384 threads): w0 = TMA/bulk producer, w1 = MMA issuer, w2 = TMEM allocator,
// w4-7 = INT4->INT8 unpack (+ level-2 q*s2+z for per-group) into the TMEM A ring,
// w8-11 = epilogue (TMEM -> regs -> fp32 math -> fp16 -> smem transpose -> 16 B global stores).
//
// Scheduling: persistent.  "DP" mode walks whole output tiles (prefill);  "SK" (stream-K) mode gives
// every CTA an equal contiguous range of K-blocks so that small-M (decode) problems fill all SMs; tiles
// that span CTAs are summed exactly in INT32 with cp.reduce.async.bulk (.add.s32) into an L2-resident
// workspace and finished by the last contributor.
#include "launch.h"
#include "norm_tail.cuh"
#include "ptx.cuh"
#include "w4a8_gemm.h"

#include <algorithm>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <unordered_map>

namespace ob {

constexpr int BM = 128;           // weight rows per tile  (UMMA M)
constexpr int BK = 128;           // K bytes per pipeline stage (= one level-2 group)
constexpr int W_STAGE = BM * BK / 2;  // 8192 packed bytes
constexpr int NUM_THREADS = 384;
constexpr int A_COLS_PER_STAGE = BK / 4;  // 32 TMEM columns hold 128 x 128 int8

template <int BN, bool TWO = false>
struct Cfg {
  // Three decoupled rings.  Packed weights come from HBM (long latency, nothing downstream holds them once they
  // are unpacked): deep ring, released by the unpack warps.  Activations come from L2 and the unpacked INT8
  // weights live in TMEM: shallow rings, released when the MMAs that read them retire.
  // TWO: CTA-pair MMA (tcgen05 cta_group::2, 256 weight rows x BN tokens per pair): each CTA stages only BN/2 token
  // rows of the activation tile, so the bytes entering an SM per K-block drop from 8 KB + BN*128 to 8 KB + BN*64.
  static constexpr int W_STAGES = TWO ? 8 : ((BN >= 128) ? 6 : (BN >= 64 ? 12 : 16));
  static constexpr int B_STAGES = TWO ? 8 : ((BN >= 128) ? 6 : 8);   // activation ring and TMEM A ring advance in
  static constexpr int A_SLOTS = B_STAGES;                             // lock-step: one commit per K-block frees both
  static constexpr int B_ROWS = TWO ? BN / 2 : BN;       // token rows of the tile held by this CTA
  static constexpr int B_STAGE = B_ROWS * BK;
  static constexpr int S2_STAGE = 256;
  static constexpr int ACC_BUFS = 2;
  static constexpr int TMEM_A_BASE = ACC_BUFS * BN;  // columns
  static constexpr int TMEM_COLS = 512;
  static constexpr int OUT_PITCH = BM + 8;            // halves
  static constexpr int STAGING = (BN * OUT_PITCH * 2 > (BN + 8) * BM * 4) ? BN * OUT_PITCH * 2 : (BN + 8) * BM * 4;
  static constexpr int SMEM_B = 0;
  static constexpr int SMEM_W = SMEM_B + B_STAGES * B_STAGE;
  static constexpr int SMEM_S2 = SMEM_W + W_STAGES * W_STAGE;
  static constexpr int SMEM_STAGING = SMEM_S2 + W_STAGES * S2_STAGE;
  static constexpr int SMEM_TOK = SMEM_STAGING + STAGING;      // float sa[BN], ss[BN]
  static constexpr int SMEM_BAR = SMEM_TOK + BN * 8;
  static constexpr int NUM_BARS = 2 * W_STAGES + 4 * B_STAGES + 2 * ACC_BUFS;
  static constexpr int SMEM_MISC = SMEM_BAR + NUM_BARS * 8;    // tmem slot, flags
  static constexpr int SMEM_TOTAL = SMEM_MISC + 64 + 1024;      // + alignment slack
  static_assert(TMEM_A_BASE + A_SLOTS * A_COLS_PER_STAGE <= 512, "TMEM budget");
  static_assert(SMEM_TOTAL <= 227 * 1024, "smem budget");
};

struct GemmParams {
  const int8_t* qweight;    // [N/32][K/32][512]
  const int8_t* s2_scales;  // per-group: [K/128][N] (N permuted in 32-blocks), else null
  const int8_t* s2_zeros;
  const __half* wscales;    // [N]
  const __half* ascales;    // [M]
  const __half* w_szs;      // per-channel: [N]
  const __half* a_ssums;    // per-channel: [M]
  __half* out;              // [M, ldc]
  int32_t* ws;              // split-K workspace: [grid][BN*128] int32, zero between launches
  int32_t* counters;        // [grid]
  int M, N, K, ldc;
  int n_tiles, m_tiles, kb_per_tile;
  int mode;                 // 0 = DP tiles, 1 = stream-K (L2 bulk-reduce), 2 = cluster split-K (DSMEM reduce-scatter)
  int cluster_k;            // mode 2: CTAs per tile
  int mc;                   // mode 0: activation multicast across a cluster of mc N-tiles (1 = off)
  int units_per_cta;        // SK: K-blocks per CTA
  int group_m;              // DP raster: m-tiles per L2 group
  // Optional tail (extension): once every tile of this GEMM is written, the epilogue warps of CTA r run the add + norm +
  // quant(+sum) that consumes output row r (llama_w4a8_unpad.py:425-431,437 + next layer's :416-421) -- one launch and
  // one grid-wide barrier instead of a kernel boundary.  tail = 0 off, 1 on.
  int tail;
  const __half* t_hidden_in; __half* t_hidden_out; const __half* t_gamma;
  int8_t* t_q; __half* t_scale; __half* t_sum; float t_eps;
  unsigned int* t_counter; unsigned int* t_gen;   // grid barrier (sense reversal), in the split-K counter workspace
  long long* dbg_t;         // timing experiments only (OB_GEMM_DBGT = device address): [grid][16] wait cycles per role
  int unpack2;              // unpack warps take two K-blocks per iteration when both have landed (OB_GEMM_UNPACK2, default 1)
  int w_rows2k;             // weight tensor map variant: 4 rows of 2 KB per K-block instead of 16 rows of 512 B
  int dbg;                  // timing experiments only (OB_GEMM_DBG): 1 = no wait::st, 2 = no unpack, 4 = no MMA
};

struct Seg {
  int tile, kb0, kb1;
};

struct SegIter {
  int mode, KB, total_tiles, pos, end, tile, step;
  OB_DEVICE void init(const GemmParams& p) {
    mode = p.mode;
    KB = p.kb_per_tile;
    total_tiles = p.n_tiles * p.m_tiles;
    if (mode == 2) {
      // one segment: tile = cluster index, K-range = rank-th slice of the tile's K-blocks
      const int k = p.cluster_k, r = (int)cluster_ctarank();
      tile = (int)cluster_id_x();
      pos = (int)(((long long)KB * r) / k);
      end = (int)(((long long)KB * (r + 1)) / k);
      step = 0;
    } else if (mode == 0) {
      // with multicast the unit of scheduling is a super-tile (mc N-tiles x one token tile) per cluster
      tile = blockIdx.x / p.mc;
      step = gridDim.x / p.mc;
      total_tiles = (p.n_tiles / p.mc) * p.m_tiles;
    } else {
      long long tot = (long long)total_tiles * KB;
      long long b = (long long)blockIdx.x * p.units_per_cta;
      long long e = b + p.units_per_cta;
      pos = (int)(b < tot ? b : tot);
      end = (int)(e < tot ? e : tot);
    }
  }
  OB_DEVICE bool next(Seg& s) {
    if (mode == 2) {
      if (step) return false;
      step = 1;
      s.tile = tile; s.kb0 = pos; s.kb1 = end;
      return end > pos;
    }
    if (mode == 0) {
      if (tile >= total_tiles) return false;
      s.tile = tile;
      s.kb0 = 0;
      s.kb1 = KB;
      tile += step;
      return true;
    }
    if (pos >= end) return false;
    s.tile = pos / KB;
    s.kb0 = pos - s.tile * KB;
    int room = KB - s.kb0;
    int left = end - pos;
    s.kb1 = s.kb0 + (left < room ? left : room);
    pos += s.kb1 - s.kb0;
    return true;
  }
};

OB_DEVICE void tile_coords(const struct GemmParams& p, int tile, int& nt, int& mt);

// Flat iteration over this CTA's K-blocks (tile, kb); tile coordinates are recomputed only per segment.
struct KbIter {
  SegIter it;
  Seg sg;
  int kb, nt, mt;
  OB_DEVICE void init(const GemmParams& p) { it.init(p); sg.kb0 = sg.kb1 = 0; kb = 0; nt = mt = 0; }
  OB_DEVICE bool next(const GemmParams& p) {
    if (kb + 1 < sg.kb1 && sg.kb1 > sg.kb0) { ++kb; return true; }
    if (!it.next(sg)) return false;
    kb = sg.kb0;
    tile_coords(p, sg.tile, nt, mt);
    return true;
  }
};

OB_DEVICE void tile_coords(const GemmParams& p, int tile, int& nt, int& mt) {
  if (p.mode == 0) {
    const int n_super = p.n_tiles / p.mc;
    int per_group = p.group_m * n_super;
    int g = tile / per_group;
    int r = tile - g * per_group;
    int gm = min(p.group_m, p.m_tiles - g * p.group_m);
    nt = r / gm;
    mt = g * p.group_m + (r - nt * gm);
    if (p.mc > 1) nt = nt * p.mc + (int)cluster_ctarank();
  } else {
    nt = tile / p.m_tiles;
    mt = tile - nt * p.m_tiles;
  }
}

// Timing experiments (compile with -DOB_GEMM_TIMING, run tools/gemm_waits.py): accumulate the cycles a role spends in a
// wait into tw[slot] and dump them to p.dbg_t.  Compiled out by default: even disabled at run time the extra branches and
// accumulators cost ~20 % of the kernel's throughput (measured, round 1).
// Timing / ablation experiments (tools/gemm_micro.py): `-DOB_GEMM_DEBUG` compiles the OB_GEMM_DBG bit tests into the
// role loops; the shipped build has none of these branches.
#ifdef OB_GEMM_DEBUG
#define OB_DBG(mask) ((p.dbg & (mask)) != 0)
#else
#define OB_DBG(mask) false
#endif

#ifdef OB_GEMM_TIMING
#define OB_TW(slot, call)                  \
  do {                                     \
    if (p.dbg_t) {                         \
      const long long _t0 = clock64();     \
      call;                                \
      tw[slot] += clock64() - _t0;         \
    } else {                               \
      call;                                \
    }                                      \
  } while (0)
#define OB_TW_DECL(n) long long tw[n] = {}
#define OB_TW_DUMP(stmt) do { if (p.dbg_t) { stmt; } } while (0)
#else
#define OB_TW(slot, call) call
#define OB_TW_DECL(n) (void)0
#define OB_TW_DUMP(stmt) (void)0
#endif

// Bytewise (a + b) mod 256 on four packed bytes; `__vadd4` of the reference (per_group/gemm_cuda.cu:307).
OB_DEVICE uint32_t vadd4(uint32_t a, uint32_t b) {
  uint32_t s = (a & 0x7f7f7f7fu) + (b & 0x7f7f7f7fu);
  return s ^ ((a ^ b) & 0x80808080u);
}

template <int BN, bool PER_GROUP, bool TWO>
__global__ void __launch_bounds__(NUM_THREADS, 1)
w4a8_gemm_kernel(const __grid_constant__ CUtensorMap act_map, const __grid_constant__ CUtensorMap w_map,
                 const __grid_constant__ CUtensorMap act_mc_map, const GemmParams p) {
  using C = Cfg<BN, TWO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem + C::SMEM_B;
  uint8_t* sW = smem + C::SMEM_W;
  uint8_t* sS2 = smem + C::SMEM_S2;
  uint8_t* sStage = smem + C::SMEM_STAGING;
  float* sTok = reinterpret_cast<float*>(smem + C::SMEM_TOK);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::SMEM_BAR);
  uint64_t* w_full = bars;                               // packed weights (+ s2) landed
  uint64_t* w_empty = w_full + C::W_STAGES;               // unpack warps read the stage into registers
  uint64_t* b_full = w_empty + C::W_STAGES;               // activation tile landed
  uint64_t* a_full = b_full + C::B_STAGES;                // unpack warps filled the TMEM A slot
  uint64_t* ba_empty = a_full + C::B_STAGES;              // MMAs that read activation stage s / TMEM slot s retired
  uint64_t* b_empty_mc = ba_empty + C::B_STAGES;          // multicast only: every CTA of the cluster released stage s
  uint64_t* acc_full = b_empty_mc + C::B_STAGES;
  uint64_t* acc_empty = acc_full + C::ACC_BUFS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::SMEM_MISC);
  int* sFlag = reinterpret_cast<int*>(smem + C::SMEM_MISC + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&act_map);
      tma_prefetch_desc(&w_map);
    }
    // one barrier per lane: arrival counts are 4 for the barriers the four unpack / epilogue warps arrive on
    for (int i = lane; i < C::NUM_BARS; i += 32) {
      uint64_t* b = bars + i;
      const bool four = (b >= w_empty && b < b_full) || (b >= a_full && b < ba_empty) || (b >= acc_empty);
      const bool mcb = (b >= b_empty_mc && b < acc_full);
      uint32_t cnt = mcb ? (uint32_t)p.mc : (four ? 4u : 1u);
      if (TWO) {
        if (mcb) cnt = 1;                    // reused as peer_ready[]: the peer CTA's stage s (B half + A slot) is ready
        if (b >= acc_empty) cnt = 8;         // leader: the epilogue warps of both CTAs drained the accumulator
      }
      mbar_init(b, cnt);
    }
    mbar_fence_init();
  }
  if (warp == 2) { if (TWO) tmem_alloc2<C::TMEM_COLS>(tmem_slot); else tmem_alloc<C::TMEM_COLS>(tmem_slot); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // all CTAs of the cluster run (barriers initialised) before any DSMEM store / multicast copy / remote arrive targets them
  if ((p.mode == 2 && !OB_DBG(16)) || p.mc > 1) cluster_barrier();
  const uint32_t tmem_base = *tmem_slot;


  if (warp == 0) {
    // ================================================================ weight producer
    // One 3-D TMA per stage brings the 4 (n32) x 4 (k32) x 512 B packed-weight blocks of the tile's K-block
    // (the reference layout [N/32][K/32][512 B] viewed as a tensor of 8-byte elements).  Weights never depend on
    // the previous kernel, so under programmatic dependent launch they stream while it is still draining.
    if (lane == 0) {
      KbIter it;
      it.init(p);
      int stage = 0, phase = 0;
      OB_TW_DECL(1);
#ifdef OB_GEMM_TIMING
      const long long t_start = clock64();
#endif
      while (it.next(p)) {
        OB_TW(0, mbar_wait(&w_empty[stage], phase ^ 1));
        const int n_cnt = min(BM, p.N - it.nt * BM);
        mbar_arrive_expect_tx(&w_full[stage], (OB_DBG(64) ? 0 : W_STAGE) + (PER_GROUP ? 2 * n_cnt : 0));
        if (!OB_DBG(64)) {
          if (p.w_rows2k) tma_load_3d(sW + stage * W_STAGE, &w_map, 0, it.kb, it.nt * 4, &w_full[stage]);
          else tma_load_3d(sW + stage * W_STAGE, &w_map, 0, it.kb * 4, it.nt * 4, &w_full[stage]);
        }
        if (PER_GROUP) {
          bulk_g2s(sS2 + stage * C::S2_STAGE, p.s2_scales + (size_t)it.kb * p.N + it.nt * BM, n_cnt, &w_full[stage]);
          bulk_g2s(sS2 + stage * C::S2_STAGE + 128, p.s2_zeros + (size_t)it.kb * p.N + it.nt * BM, n_cnt, &w_full[stage]);
        }
        if (++stage == C::W_STAGES) { stage = 0; phase ^= 1; }
      }
      OB_TW_DUMP(p.dbg_t[blockIdx.x * 16 + 0] = tw[0]; p.dbg_t[blockIdx.x * 16 + 9] = clock64() - t_start);
    }
  } else if (warp == 3) {
    // ================================================================ activation producer (own warp: TMA issue
    // is slow per thread, two issuers double the rate); these tiles are the previous kernel's output.
    if (lane == 0) {
      pdl_wait();
      KbIter it;
      it.init(p);
      int stage = 0, phase = 0;
      OB_TW_DECL(1);
      if (TWO) {
        // CTA pair: this CTA stages only its half of the token rows; the pair MMA reads both halves.
        const int rank = (int)cluster_ctarank();
        while (it.next(p)) {
          OB_TW(0, mbar_wait(&ba_empty[stage], phase ^ 1));
          mbar_arrive_expect_tx(&b_full[stage], C::B_STAGE);
          tma_load_2d(sB + stage * C::B_STAGE, &act_mc_map, it.kb * BK, it.mt * BN + rank * C::B_ROWS, &b_full[stage]);
          if (++stage == C::B_STAGES) { stage = 0; phase ^= 1; }
        }
      } else if (p.mc > 1) {
        // Multicast: this CTA fetches rows [rank*BN/mc, (rank+1)*BN/mc) of the token tile and delivers them to every
        // CTA of the cluster (they work on different weight rows, same tokens, same K-block): L2 is read once per
        // cluster instead of once per CTA.  A stage may be refilled only when all mc CTAs have retired its MMAs.
        const int rank = (int)cluster_ctarank();
        const int rows = BN / p.mc;
        const uint16_t mask = (uint16_t)((1u << p.mc) - 1u);
        while (it.next(p)) {
          mbar_wait(&b_empty_mc[stage], phase ^ 1);
          mbar_arrive_expect_tx(&b_full[stage], C::B_STAGE);
          tma_load_2d_mc(sB + stage * C::B_STAGE + rank * rows * BK, &act_mc_map, it.kb * BK, it.mt * BN + rank * rows,
                         &b_full[stage], mask);
          if (++stage == C::B_STAGES) { stage = 0; phase ^= 1; }
        }
      } else {
        while (it.next(p)) {
          OB_TW(0, mbar_wait(&ba_empty[stage], phase ^ 1));
          mbar_arrive_expect_tx(&b_full[stage], OB_DBG(32) ? 0 : C::B_STAGE);
          if (!OB_DBG(32)) tma_load_2d(sB + stage * C::B_STAGE, &act_map, it.kb * BK, it.mt * BN, &b_full[stage]);
          if (++stage == C::B_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      OB_TW_DUMP(p.dbg_t[blockIdx.x * 16 + 1] = tw[0]);
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    SegIter it;
    it.init(p);
    Seg sg;
    // The issue loop is the throughput limiter for narrow tiles (4 short MMAs per K-block), so it is kept lean:
    // one elected lane (elect.sync lets ptxas use plain uniform-register moves instead of a waterfall loop),
    // descriptors advanced by adds from loop-invariant bases, one commit per K-block.
    int st = 0, ph = 0, acc = 0, acc_phase = 0;
    constexpr uint32_t idesc = umma_idesc_i8(TWO ? 2 * BM : BM, BN, true, true);
    const uint64_t bdesc0 = umma_desc_kmajor_sw128(smem_u32(sB));
    const uint32_t a_tmem0 = tmem_base + C::TMEM_A_BASE;
    OB_TW_DECL(6);
    if (TWO) {
      uint64_t* peer_ready = b_empty_mc;
      if (cluster_ctarank() == 0) {
        // leader: issues the pair MMAs once its own stage and the peer's stage are both ready
        while (it.next(sg)) {
          OB_TW(2, mbar_wait(&acc_empty[acc], acc_phase ^ 1));
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
            OB_TW(3, mbar_wait(&b_full[st], ph));
            OB_TW(4, mbar_wait(&a_full[st], ph));
            OB_TW(5, mbar_wait(&peer_ready[st], ph));
            tc_fence_after();
            if (elect_one()) {
              const uint64_t bdesc = bdesc0 + (uint64_t)((st * C::B_STAGE) >> 4);
              const uint32_t a_tmem = a_tmem0 + st * A_COLS_PER_STAGE;
              umma2_i8_ts(d_tmem, a_tmem, bdesc, idesc, kb > sg.kb0 ? 1u : 0u);
              umma2_i8_ts(d_tmem, a_tmem + 8, bdesc + 2, idesc, 1u);
              umma2_i8_ts(d_tmem, a_tmem + 16, bdesc + 4, idesc, 1u);
              umma2_i8_ts(d_tmem, a_tmem + 24, bdesc + 6, idesc, 1u);
              umma2_commit_mc(&ba_empty[st], 3);                       // frees stage s in both CTAs
              if (kb == sg.kb1 - 1) umma2_commit_mc(&acc_full[acc], 3);  // both epilogues may drain
            }
            __syncwarp();
            if (++st == C::B_STAGES) { st = 0; ph ^= 1; }
          }
          if (++acc == C::ACC_BUFS) { acc = 0; acc_phase ^= 1; }
        }
      } else {
        // peer: forwards "my activation half landed and my TMEM A slot is filled" to the leader
        while (it.next(sg)) {
          for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
            OB_TW(3, mbar_wait(&b_full[st], ph));
            OB_TW(4, mbar_wait(&a_full[st], ph));
            tc_fence_after();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive_remote(&peer_ready[st], 0);
            __syncwarp();
            if (++st == C::B_STAGES) { st = 0; ph ^= 1; }
          }
        }
      }
    } else
    while (it.next(sg)) {
      OB_TW(2, mbar_wait(&acc_empty[acc], acc_phase ^ 1));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
        OB_TW(3, mbar_wait(&b_full[st], ph));
        OB_TW(4, mbar_wait(&a_full[st], ph));
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bdesc = bdesc0 + (uint64_t)((st * C::B_STAGE) >> 4);
          const uint32_t a_tmem = a_tmem0 + st * A_COLS_PER_STAGE;
          if (!OB_DBG(4)) {
            umma_i8_ts(d_tmem, a_tmem, bdesc, idesc, kb > sg.kb0 ? 1u : 0u);
            umma_i8_ts(d_tmem, a_tmem + 8, bdesc + 2, idesc, 1u);
            umma_i8_ts(d_tmem, a_tmem + 16, bdesc + 4, idesc, 1u);
            umma_i8_ts(d_tmem, a_tmem + 24, bdesc + 6, idesc, 1u);
          }
          umma_commit(&ba_empty[st]);
          if (p.mc > 1) umma_commit_mc(&b_empty_mc[st], (uint16_t)((1u << p.mc) - 1u));
          if (kb == sg.kb1 - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++st == C::B_STAGES) { st = 0; ph ^= 1; }
      }
      if (++acc == C::ACC_BUFS) { acc = 0; acc_phase ^= 1; }
    }
    OB_TW_DUMP(if (lane == 0) for (int i = 2; i < 6; ++i) p.dbg_t[blockIdx.x * 16 + i] = tw[i]);
  } else if (warp >= 4 && warp < 8) {
    // ================================================================ INT4 -> INT8 unpack into TMEM
    const int q = warp - 4;  // TMEM lane quarter == n32 block inside the tile
    SegIter it;
    it.init(p);
    Seg sg;
    // Software pipelined: the TMEM stores of K-block i are only waited for (tcgen05.wait::st) while the packed
    // bytes of the next K-block(s) are already on their way from shared memory, so the store-completion latency is off
    // the per-K-block critical path of this warp.  a_full(i) is therefore signalled one iteration late -- the TMEM ring
    // absorbs that -- and once more after the last K-block.
    // Two K-blocks per iteration when the second one has already landed (never waited for): the role is a latency
    // chain (mbarrier wait -> ld.shared -> tcgen05.st -> arrive; ~78 % busy with ~110 instructions per K-block,
    // profiles/r1_gemm_role_waits.log), and pairing K-blocks halves the number of chain traversals.
    int ws = 0, wph = 0, as = 0, aph = 0, pend0 = -1, pend1 = -1;
    OB_TW_DECL(2);
    const uint32_t sW_u32 = smem_u32(sW), sS2_u32 = smem_u32(sS2);
    const uint32_t t_q = tmem_base + ((uint32_t)(q * 32) << 16) + C::TMEM_A_BASE;

    auto load_stage = [&](int st, uint4 (&v)[4]) {
      const uint32_t wsm = sW_u32 + st * W_STAGE + q * 2048 + lane * 16;
#pragma unroll
      for (int a = 0; a < 4; ++a) v[a] = lds_v4(wsm + a * 512);
    };
    auto convert_store = [&](int st, int slot, const uint4 (&v)[4]) {
      uint32_t sc[4], zr[4];
      if (PER_GROUP) {
        const uint32_t ps = lds_u32(sS2_u32 + st * C::S2_STAGE + q * 32 + (lane >> 2) * 4);
        const uint32_t pz = lds_u32(sS2_u32 + st * C::S2_STAGE + 128 + q * 32 + (lane >> 2) * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sc[j] = (ps >> (8 * j)) & 0xFFu;
          zr[j] = ((pz >> (8 * j)) & 0xFFu) * 0x01010101u;
        }
      }
      const uint32_t t_lo = t_q + slot * A_COLS_PER_STAGE;
      const uint32_t t_hi = t_lo + (16u << 16);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        uint32_t l0 = v[a].x & 0x0F0F0F0Fu, l1 = v[a].y & 0x0F0F0F0Fu, l2 = v[a].z & 0x0F0F0F0Fu, l3 = v[a].w & 0x0F0F0F0Fu;
        uint32_t h0 = (v[a].x >> 4) & 0x0F0F0F0Fu, h1 = (v[a].y >> 4) & 0x0F0F0F0Fu, h2 = (v[a].z >> 4) & 0x0F0F0F0Fu,
                 h3 = (v[a].w >> 4) & 0x0F0F0F0Fu;
        if (PER_GROUP) {
          // rows: l0,l2 -> c (scale 0); l1,l3 -> c+8 (scale 1); h0,h2 -> c+16 (scale 2); h1,h3 -> c+24 (scale 3)
          l0 = vadd4(l0 * sc[0], zr[0]); l2 = vadd4(l2 * sc[0], zr[0]);
          l1 = vadd4(l1 * sc[1], zr[1]); l3 = vadd4(l3 * sc[1], zr[1]);
          h0 = vadd4(h0 * sc[2], zr[2]); h2 = vadd4(h2 * sc[2], zr[2]);
          h1 = vadd4(h1 * sc[3], zr[3]); h3 = vadd4(h3 * sc[3], zr[3]);
        }
        tmem_st_16x128b_x2(t_lo + a * 8, l0, l1, l2, l3);
        tmem_st_16x128b_x2(t_hi + a * 8, h0, h1, h2, h3);
      }
    };
    auto flush_pending = [&]() {
      if (pend0 >= 0) {   // previous K-block(s): their TMEM stores have had a whole iteration to land
        if (!OB_DBG(1)) tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a_full[pend0]);
          if (pend1 >= 0) mbar_arrive(&a_full[pend1]);
        }
        pend0 = pend1 = -1;
      }
    };

    while (it.next(sg)) {
      for (int kb = sg.kb0; kb < sg.kb1;) {
        OB_TW(0, mbar_wait(&w_full[ws], wph));
        if (OB_DBG(2)) {
          __syncwarp();
          if (lane == 0) { mbar_arrive(&w_empty[ws]); }
          mbar_wait(&ba_empty[as], aph ^ 1);
          if (lane == 0) { mbar_arrive(&a_full[as]); }
          if (++ws == C::W_STAGES) { ws = 0; wph ^= 1; }
          if (++as == C::A_SLOTS) { as = 0; aph ^= 1; }
          ++kb;
          continue;
        }
        int ws1 = ws + 1, wph1 = wph, as1 = as + 1, aph1 = aph;
        if (ws1 == C::W_STAGES) { ws1 = 0; wph1 ^= 1; }
        if (as1 == C::A_SLOTS) { as1 = 0; aph1 ^= 1; }
        uint4 v0[4], v1[4];
        load_stage(ws, v0);
        // second K-block of the pair only if its bytes are already there (probe once, never wait)
        bool two = false;
        if (p.unpack2 && kb + 1 < sg.kb1) {
          const int ready = (lane == 0) ? (int)mbar_test_wait(&w_full[ws1], wph1) : 0;
          two = __shfl_sync(0xffffffffu, ready, 0) != 0;
        }
        if (two) load_stage(ws1, v1);
        flush_pending();
        OB_TW(1, mbar_wait(&ba_empty[as], aph ^ 1));
        if (two) mbar_wait(&ba_empty[as1], aph1 ^ 1);
        tc_fence_after();
        convert_store(ws, as, v0);
        if (two) convert_store(ws1, as1, v1);
        // the packed stage(s) are in registers: hand them back to the producer right away
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&w_empty[ws]);
          if (two) mbar_arrive(&w_empty[ws1]);
        }
        pend0 = as;
        pend1 = two ? as1 : -1;
        if (two) {
          ws = ws1; wph = wph1; as = as1; aph = aph1;
        }
        if (++ws == C::W_STAGES) { ws = 0; wph ^= 1; }
        if (++as == C::A_SLOTS) { as = 0; aph ^= 1; }
        kb += two ? 2 : 1;
      }
    }
    flush_pending();
    OB_TW_DUMP(if (warp == 4 && lane == 0) { p.dbg_t[blockIdx.x * 16 + 6] = tw[0]; p.dbg_t[blockIdx.x * 16 + 7] = tw[1]; });
  } else if (warp >= 8) {
    // ================================================================ epilogue
    const int q = warp - 8;
    const int et = threadIdx.x - 256;  // 0..127
    SegIter it;
    it.init(p);
    Seg sg;
    int acc = 0, acc_phase = 0;
    __half* stage16 = reinterpret_cast<__half*>(sStage);
    int32_t* stage32 = reinterpret_cast<int32_t*>(sStage);
    bool cluster_done = false;
    OB_TW_DECL(1);
    pdl_wait();  // ascales / a_ssums come from the previous kernel; `out` may still be read by it
    while (it.next(sg)) {
      int nt, mt;
      tile_coords(p, sg.tile, nt, mt);
      const int m0 = mt * BN;
      const int n_row = nt * BM + q * 32 + lane;
      const bool full_tile = (sg.kb0 == 0 && sg.kb1 == p.kb_per_tile) && p.mode != 2;
      const bool n_ok = n_row < p.N;
      float wsc = 0.f, wsz = 0.f;
      if (n_ok) {
        wsc = __half2float(p.wscales[n_row]);
        if (!PER_GROUP) wsz = __half2float(p.w_szs[n_row]);
      }
      // previous segment's staging users are done (barrier at the end of the loop body)
      if (et < BN) {
        const int m = m0 + et;
        sTok[et] = (m < p.M) ? __half2float(p.ascales[m]) : 0.f;
        sTok[BN + et] = (!PER_GROUP && m < p.M) ? __half2float(p.a_ssums[m]) : 0.f;
      }
      OB_TW(0, mbar_wait(&acc_full[acc], acc_phase));
      tc_fence_after();
      asm volatile("bar.sync 1, 128;" ::: "memory");  // sTok visible
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      if (full_tile) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(t_acc + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float ps = __int2float_rn((int)r[j]);
            float o;
            if (PER_GROUP) {
              o = ps * (wsc * sTok[c0 + j]);
            } else {
              o = __fmaf_rn(-wsz, sTok[BN + c0 + j], (ps * wsc) * sTok[c0 + j]);
            }
            stage16[(c0 + j) * C::OUT_PITCH + q * 32 + lane] = __float2half_rn(o);
          }
        }
      } else if (p.mode == 2) {
        // cluster split-K: reduce-scatter of the INT32 partials over distributed shared memory.  Rank o owns the
        // token columns [o*cp, (o+1)*cp); every CTA deposits its partial for those columns in slot [its rank] of
        // o's staging buffer (st.shared::cluster), exact and order independent.
        const int k = p.cluster_k, my = (int)cluster_ctarank();
        const int cp = (BN + k - 1) / k;
        const uint32_t base_local = smem_u32(stage32);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(t_acc + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = c0 + j;
            const int owner = col / cp;
            const int lc = col - owner * cp;
            const uint32_t off = (uint32_t)(((my * cp + lc) * BM + q * 32 + lane) * 4);
            if (owner == my || OB_DBG(8)) stage32[(my * cp + lc) * BM + q * 32 + lane] = (int)r[j];
            else st_shared_cluster_u32(mapa_shared(base_local + off, (uint32_t)owner), r[j]);
          }
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(t_acc + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) stage32[(c0 + j) * BM + q * 32 + lane] = (int)r[j];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (TWO && cluster_ctarank() != 0) mbar_arrive_remote(&acc_empty[acc], 0);
        else mbar_arrive(&acc_empty[acc]);
      }
      if (++acc == C::ACC_BUFS) { acc = 0; acc_phase ^= 1; }

      if (full_tile) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // coalesced 16 B stores: 16 chunks of 8 halves per token row
        const int chunk = et & 15;
        const int n_base = nt * BM + chunk * 8;
        if (n_base < p.N) {
          for (int row = et >> 4; row < BN; row += 8) {
            const int m = m0 + row;
            if (m < p.M) {
              const uint4 v = *reinterpret_cast<const uint4*>(stage16 + row * C::OUT_PITCH + chunk * 8);
              *reinterpret_cast<uint4*>(p.out + (size_t)m * p.ldc + n_base) = v;
            }
          }
        }
      } else if (p.mode == 2) {
        const int k = p.cluster_k, my = (int)cluster_ctarank();
        const int cp = (BN + k - 1) / k;
        cluster_barrier();  // every partial of this tile has landed in its owner's shared memory
        cluster_done = true;
        const int ncols = min(cp, BN - my * cp);
        const int row = q * 32 + lane;
        for (int lc = 0; lc < ncols; ++lc) {
          int sum = 0;
          for (int src = 0; src < k; ++src) sum += stage32[(src * cp + lc) * BM + row];
          const int col = my * cp + lc;
          const int m = m0 + col;
          if (m < p.M && n_ok) {
            const float ps = __int2float_rn(sum);
            float o;
            if (PER_GROUP) o = ps * (wsc * sTok[col]);
            else o = __fmaf_rn(-wsz, sTok[BN + col], (ps * wsc) * sTok[col]);
            p.out[(size_t)m * p.ldc + n_row] = __float2half_rn(o);
          }
        }
      } else {
        // exact INT32 split-K: bulk reduce-add the partial tile into the L2 workspace slot of this tile
        const int KB = p.kb_per_tile;
        const int first_cta = (int)(((long long)sg.tile * KB) / p.units_per_cta);
        const int last_cta = (int)(((long long)(sg.tile + 1) * KB - 1) / p.units_per_cta);
        const int contributors = last_cta - first_cta + 1;
        int32_t* slot = p.ws + (size_t)first_cta * (BN * BM);
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          bulk_reduce_add_s32(slot, stage32, BN * BM * 4);
          bulk_commit();
          bulk_wait_all();
          asm volatile("fence.proxy.async.global;" ::: "memory");
          __threadfence();
          const int old = atomicAdd(&p.counters[first_cta], 1);
          *sFlag = (old == contributors - 1) ? 1 : 0;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (*sFlag) {
          __threadfence();
          // finalize: elementwise over [BN tokens][128 n]; thread et owns the 8 columns n_base.. of rows
          // (et >> 4) + 8 i.  All L2 reads of a batch are issued before anything depends on them (the slot was
          // written by other SMs, so every read is an L2 round trip: batching turns BN/8 dependent trips into one).
          const int chunk = et & 15;
          const int n_base = nt * BM + chunk * 8;
          const bool n_in = n_base < p.N;
          float wsc8[8], wsz8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            wsc8[j] = n_in ? __half2float(p.wscales[n_base + j]) : 0.f;
            wsz8[j] = (!PER_GROUP && n_in) ? __half2float(p.w_szs[n_base + j]) : 0.f;
          }
          constexpr int ROWS_PER_BATCH = 8;   // 8 rows x 2 int4 = 64 registers in flight
#pragma unroll 1
          for (int r0 = et >> 4; r0 < BN; r0 += 8 * ROWS_PER_BATCH) {
            int4 a0[ROWS_PER_BATCH], a1[ROWS_PER_BATCH];
#pragma unroll
            for (int i = 0; i < ROWS_PER_BATCH; ++i) {
              const int row = r0 + 8 * i;
              if (row < BN) {
                const int4* src = reinterpret_cast<const int4*>(slot + row * BM + chunk * 8);
                a0[i] = __ldcg(src);
                a1[i] = __ldcg(src + 1);
              }
            }
#pragma unroll
            for (int i = 0; i < ROWS_PER_BATCH; ++i) {
              const int row = r0 + 8 * i;
              if (row < BN) {
                int4* src = reinterpret_cast<int4*>(slot + row * BM + chunk * 8);
                __stcg(src, make_int4(0, 0, 0, 0));
                __stcg(src + 1, make_int4(0, 0, 0, 0));
                const int m = m0 + row;
                if (m < p.M && n_in) {
                  const int av[8] = {a0[i].x, a0[i].y, a0[i].z, a0[i].w, a1[i].x, a1[i].y, a1[i].z, a1[i].w};
                  const float sa = sTok[row], ss = sTok[BN + row];
                  __align__(16) __half o[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float ps = __int2float_rn(av[j]);
                    float r;
                    if (PER_GROUP) r = ps * (wsc8[j] * sa);
                    else r = __fmaf_rn(-wsz8[j], ss, (ps * wsc8[j]) * sa);
                    o[j] = __float2half_rn(r);
                  }
                  *reinterpret_cast<uint4*>(p.out + (size_t)m * p.ldc + n_base) = *reinterpret_cast<uint4*>(o);
                }
              }
            }
          }
          if (et == 0) p.counters[first_cta] = 0;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // staging + sTok free for the next segment
    }
    OB_TW_DUMP(if (et == 0) p.dbg_t[blockIdx.x * 16 + 8] = tw[0]);
    if (p.tail) {
      // ---- grid-wide barrier among the epilogue warps (all CTAs are resident: grid <= #SMs, one CTA per SM), then
      // ---- row r of the add+norm+quant by CTA r.  Other warps of the CTA are idle by now.
      __shared__ float tail_red[64];
      __threadfence();                                     // this CTA's output tiles are visible device-wide
      tail_bar();
      if (et == 0) {
        unsigned int gen;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(p.t_gen) : "memory");
        if (atomicAdd(p.t_counter, 1u) == gridDim.x - 1) {
          *p.t_counter = 0u;
          __threadfence();
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.t_gen) : "memory");
        } else {
          unsigned int g2;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g2) : "l"(p.t_gen) : "memory");
          } while (g2 == gen);
        }
        __threadfence();
      }
      tail_bar();
      for (int row = blockIdx.x; row < p.M; row += gridDim.x) {
        const __half* drow = p.out + (size_t)row * p.ldc;
        if (p.t_sum)
          add_norm_quant_row<true>(et, p.t_hidden_in + (size_t)row * p.N, drow, p.t_hidden_out + (size_t)row * p.N, p.t_gamma,
                                   p.t_q + (size_t)row * p.N, p.t_scale + row, p.t_sum + row, p.N, p.t_eps, tail_red);
        else
          add_norm_quant_row<false>(et, p.t_hidden_in + (size_t)row * p.N, drow, p.t_hidden_out + (size_t)row * p.N, p.t_gamma,
                                    p.t_q + (size_t)row * p.N, p.t_scale + row, nullptr, p.N, p.t_eps, tail_red);
      }
    }
    if (p.mode == 2 && !cluster_done) cluster_barrier();
  }
  if (p.mode == 2 && warp < 8) cluster_barrier();  // non-epilogue warps: every thread of the cluster arrives once
  __syncwarp();
  if (p.mc > 1) cluster_barrier();  // peers may still multicast-arrive on this CTA's barriers until they are done too

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { if (TWO) tmem_dealloc2<C::TMEM_COLS>(tmem_base); else tmem_dealloc<C::TMEM_COLS>(tmem_base); }
}

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess) fn = (PFN_encodeTiled)f;
  });
  return fn;
}

struct MapKey {
  const void* ptr; int M, K, BN;
  bool operator==(const MapKey& o) const { return ptr == o.ptr && M == o.M && K == o.K && BN == o.BN; }
};
struct MapHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ (size_t)k.M * 1315423911u ^ (size_t)k.K * 2654435761u ^ (size_t)k.BN;
  }
};

int make_act_map(CUtensorMap* out, const void* ptr, int M, int K, int BN) {
  static std::unordered_map<MapKey, CUtensorMap, MapHash> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  MapKey key{ptr, M, K, BN};
  auto f = cache.find(key);
  if (f != cache.end()) { *out = f->second; return 0; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return OB_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)K};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return OB_ERR_DRIVER;
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return 0;
}

int make_w_map(CUtensorMap* out, const void* ptr, int N, int K, bool rows2k) {
  static std::unordered_map<MapKey, CUtensorMap, MapHash> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  MapKey key{ptr, N, K, rows2k ? -2 : -1};
  auto f = cache.find(key);
  if (f != cache.end()) { *out = f->second; return 0; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return OB_ERR_DRIVER;
  // [N/32][K/32][512 B] as 8-byte elements: dims (fastest first) {64, K/32, N/32}
  // rows2k: the 4 k32 blocks of a K-block are contiguous (2 KB), so the same bytes can be described as
  // {256, K/128, N/32} with a {256, 1, 4} box: 4 rows of 2 KB instead of 16 rows of 512 B (same smem image).
  cuuint64_t dims[3] = {rows2k ? 256u : 64u, (cuuint64_t)(rows2k ? K / 128 : K / 32), (cuuint64_t)(N / 32)};
  cuuint64_t strides[2] = {rows2k ? 2048u : 512u, (cuuint64_t)(K / 32) * 512};
  cuuint32_t box[3] = {rows2k ? 256u : 64u, rows2k ? 1u : 4u, 4};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return OB_ERR_DRIVER;
  if (cache.size() > 8192) cache.clear();
  cache[key] = *out;
  return 0;
}

// Per-device state (one process may drive several GPUs) and one split-K workspace per (device, stream): two streams
// running GEMMs concurrently must not share partial sums or arrival counters.  Workspaces are allocated once at their
// bounded maximum and never freed or moved, so pointers baked into captured CUDA graphs stay valid.
constexpr int MAX_DEV = 16;
constexpr int MAX_STREAMS_PER_DEV = 8;
constexpr size_t WS_INTS_PER_CTA = 128 * 128;  // BN(max 128 in SK mode) * BM
struct GemmWs { cudaStream_t st; int32_t* ws; int32_t* cnt; bool used; };
struct DevState {
  int sms = 0;
  GemmWs ws[MAX_STREAMS_PER_DEV] = {};
  int n_ws = 0;
};
static DevState g_dev[MAX_DEV];
static std::mutex g_dev_mu;

int dev_sms(int dev) {
  if (!g_dev[dev].sms) cudaDeviceGetAttribute(&g_dev[dev].sms, cudaDevAttrMultiProcessorCount, dev);
  return g_dev[dev].sms;
}

int get_workspace(int dev, cudaStream_t st, int32_t** ws, int32_t** cnt) {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DevState& d = g_dev[dev];
  if (!d.ws[0].ws) {
    // the whole pool is allocated on the first launch on this device (which must be outside graph capture: cudaMalloc is
    // not capturable); later streams only get a slot ASSIGNED, which is legal during capture
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone) return OB_ERR_ARG;
    const int sms = dev_sms(dev);
    const size_t bytes = (size_t)sms * WS_INTS_PER_CTA * 4, cbytes = (size_t)GEMM_CNT_INTS(sms) * 4;
    int32_t* all = nullptr;
    int32_t* call = nullptr;
    if (cudaMalloc(&all, bytes * MAX_STREAMS_PER_DEV) != cudaSuccess) return OB_ERR_CUDA;
    if (cudaMalloc(&call, cbytes * MAX_STREAMS_PER_DEV) != cudaSuccess) return OB_ERR_CUDA;
    cudaMemset(all, 0, bytes * MAX_STREAMS_PER_DEV);
    cudaMemset(call, 0, cbytes * MAX_STREAMS_PER_DEV);
    cudaDeviceSynchronize();
    for (int i = 0; i < MAX_STREAMS_PER_DEV; ++i) {
      d.ws[i].ws = all + (size_t)i * (bytes / 4);
      d.ws[i].cnt = call + (size_t)i * GEMM_CNT_INTS(sms);
      d.ws[i].used = false;
    }
  }
  for (int i = 0; i < d.n_ws; ++i)
    if (d.ws[i].st == st) { *ws = d.ws[i].ws; *cnt = d.ws[i].cnt; return 0; }
  if (d.n_ws == MAX_STREAMS_PER_DEV) return OB_ERR_ARG;   // more distinct GEMM streams than workspaces
  GemmWs& w = d.ws[d.n_ws++];
  w.st = st; w.used = true;
  *ws = w.ws; *cnt = w.cnt;
  return 0;
}

// Experiment switches are read from the environment ONCE (first launch); OB_GEMM_ENV_RELOAD=1 (tools/gemm_micro.py)
// re-reads them on every launch.
struct GemmEnv { int dbg, unpack2, mc, two, w2k; long long* dbg_t; bool has_two; };
static GemmEnv read_env() {
  GemmEnv e{};
  const char* v;
  v = getenv("OB_GEMM_DBG"); e.dbg = v ? atoi(v) : 0;
  v = getenv("OB_GEMM_UNPACK2"); e.unpack2 = v ? atoi(v) : 1;
  v = getenv("OB_GEMM_DBGT"); e.dbg_t = v ? reinterpret_cast<long long*>(strtoull(v, nullptr, 10)) : nullptr;
  v = getenv("OB_GEMM_MC"); e.mc = v ? atoi(v) : 1;
  v = getenv("OB_GEMM_2CTA"); e.has_two = v != nullptr; e.two = v ? atoi(v) : 0;
  v = getenv("OB_GEMM_W2K"); e.w2k = (v && atoi(v)) ? 1 : 0;
  return e;
}
static const GemmEnv& gemm_env() {
  static const bool reload = getenv("OB_GEMM_ENV_RELOAD") != nullptr;
  static GemmEnv e = read_env();
  if (reload) e = read_env();
  return e;
}

template <int BN, bool PG, bool TWO = false>
static int launch(const CUtensorMap& map, const CUtensorMap& wmap, const CUtensorMap& mcmap, GemmParams& p, int grid,
                  unsigned cluster, cudaStream_t st) {
  using C = Cfg<BN, TWO>;
  auto kern = w4a8_gemm_kernel<BN, PG, TWO>;
  static bool attr_done[MAX_DEV] = {};   // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_TOTAL) != cudaSuccess)
      return OB_ERR_CUDA;
    attr_done[dev] = true;
  }
  if (p.mc > 1) {
    // persistent multicast clusters: as many as can be co-resident (GPC sizes limit how many clusters of mc fit)
    static int max_clusters_dev[MAX_DEV][9] = {};
    int* max_clusters = max_clusters_dev[dev];
    if (!max_clusters[p.mc]) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(p.mc * 64);
      cfg.blockDim = dim3(NUM_THREADS);
      cfg.dynamicSmemBytes = C::SMEM_TOTAL;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = p.mc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = 1;
      max_clusters[p.mc] = n;
    }
    const int super_tiles = (p.n_tiles / p.mc) * p.m_tiles;
    grid = std::min(super_tiles, max_clusters[p.mc]) * p.mc;
  }
  return launch_pdl_cluster(kern, dim3(grid), dim3(NUM_THREADS), (size_t)C::SMEM_TOTAL, st, cluster, map, wmap, mcmap, p) == cudaSuccess ? 0 : OB_ERR_CUDA;
}

int w4a8_gemm_run(const W4A8GemmArgs& a, bool per_group, cudaStream_t st) {
  if (a.M <= 0) return 0;
  if (a.N % 32 != 0 || a.K % 128 != 0 || a.ldc % 8 != 0) return OB_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(a.in_feats) & 15) || (reinterpret_cast<uintptr_t>(a.out_feats) & 15) ||
      (reinterpret_cast<uintptr_t>(a.qweight) & 15))
    return OB_ERR_ALIGN;
  // decode-sized problems go to the co-resident decode kernel (w4a8_gemm_decode.cu); force_mode 0 / 1 / 2 keep this
  // file's schedules reachable (tests, A/B measurements), force_mode 3 insists on the decode kernel
  {
    static const bool v1_only = getenv("OB_GEMM_V1") != nullptr;
    if (a.M <= 64 && !a.tail_hidden_in && (a.force_mode == 3 || (a.force_mode < 0 && !v1_only))) {
      const int e = w4a8_gemm_decode_run(a, per_group, st);
      if (e != OB_ERR_SHAPE || a.force_mode == 3) return e;
    }
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= MAX_DEV) return OB_ERR_ARG;
  const int g_num_sms = dev_sms(dev);
  const int sms = (a.force_ctas > 0 && a.force_mode != 2) ? std::min(a.force_ctas, g_num_sms) : g_num_sms;
  int32_t* ws_ptr = nullptr;
  int32_t* cnt_ptr = nullptr;
  if (int e = get_workspace(dev, st, &ws_ptr, &cnt_ptr)) return e;
  const GemmEnv& env = gemm_env();

  int BN = a.M <= 16 ? 16 : a.M <= 32 ? 32 : a.M <= 64 ? 64 : 128;
  if (a.force_bn > 0) BN = a.force_bn;
  GemmParams p{};
  p.dbg = env.dbg; p.unpack2 = env.unpack2; p.dbg_t = env.dbg_t;
  p.qweight = a.qweight; p.s2_scales = a.s2_scales; p.s2_zeros = a.s2_zeros;
  p.wscales = a.wscales; p.ascales = a.ascales; p.w_szs = a.w_szs; p.a_ssums = a.a_ssums;
  p.out = a.out_feats; p.ws = ws_ptr; p.counters = cnt_ptr;
  if (a.tail_hidden_in) {
    // fused add + norm + quant tail: decode-sized M, rows of <= 4096 halves (4 vectors per tail thread), N == row length
    if (a.M > 256 || a.N > TAIL_NV * TAIL_THREADS * 8 || (a.N & 7) || !a.tail_hidden_out || !a.tail_gamma || !a.tail_q ||
        !a.tail_scale || a.force_mode == 2)
      return OB_ERR_SHAPE;
    p.tail = 1;
    p.t_hidden_in = a.tail_hidden_in; p.t_hidden_out = a.tail_hidden_out; p.t_gamma = a.tail_gamma;
    p.t_q = a.tail_q; p.t_scale = a.tail_scale; p.t_sum = a.tail_sum; p.t_eps = a.tail_eps;
    p.t_counter = reinterpret_cast<unsigned int*>(cnt_ptr + g_num_sms);
    p.t_gen = p.t_counter + 1;
  }
  p.M = a.M; p.N = a.N; p.K = a.K; p.ldc = a.ldc;
  p.n_tiles = (a.N + BM - 1) / BM;
  p.m_tiles = (a.M + BN - 1) / BN;
  p.kb_per_tile = a.K / BK;
  const long long tiles = (long long)p.n_tiles * p.m_tiles;
  int grid;
  int two = 0;
  unsigned cluster = 1;
  // Scheduling choice (cost model in K-block times, constants from tools/gemm_micro.py: the L2 bulk-reduce finalisation
  // of stream-K costs ~18 K-block times (0.27 us each)):
  //   few tiles      -> cluster split-K: k CTAs per tile (k <= 8, tiles*k <= #SMs), reduce-scatter over DSMEM
  //   medium         -> whichever of data-parallel tiles / stream-K is cheaper
  //   many tiles     -> data-parallel tiles
  int mode = a.force_mode;
  const int KB = p.kb_per_tile;
  // cluster split-K (mode 2) measured no faster than stream-K on B200 for the Llama decode shapes (launch, first-byte
  // and epilogue latencies dominate both), so it is opt-in: force_mode = 2, force_ctas = CTAs per tile.
  int k_cl = a.force_mode == 2 ? (a.force_ctas > 0 ? a.force_ctas : 4) : 1;
  k_cl = std::max(1, std::min(std::min(k_cl, 8), KB));
  if (mode == 2 && k_cl < 2) mode = 1;
  if (mode < 0) {
    if (tiles >= 8LL * sms) mode = 0;
    else {
      const long long cost_dp = ((tiles + sms - 1) / sms) * KB;
      const long long cost_sk = (tiles * KB + sms - 1) / sms + 18;
      mode = cost_sk < cost_dp ? 1 : 0;
    }
  }
  p.cluster_k = 1;
  if (mode == 2) {
    p.mode = 2;
    p.cluster_k = k_cl;
    p.units_per_cta = KB;
    cluster = (unsigned)k_cl;
    grid = (int)tiles * k_cl;
  } else if (mode == 1) {
    p.mode = 1;
    const long long total = tiles * p.kb_per_tile;
    p.units_per_cta = (int)((total + sms - 1) / sms);
    grid = (int)((total + p.units_per_cta - 1) / p.units_per_cta);
  } else {
    p.mode = 0;
    p.group_m = std::max(1, 8192 / BN);  // 8192 tokens' activations stay L2-resident while all n-tiles sweep
    p.units_per_cta = p.kb_per_tile;
    grid = (int)std::min<long long>(tiles, sms);
    // Activation multicast across a cluster of mc N-tiles (TMA .multicast::cluster) is implemented and tested, but
    // measured SLOWER than unicast on B200 for the Llama prefill shapes (M = 8192: qkv 235 vs 201 us at mc = 2, 250 at
    // mc = 4; profiles/r1_summary.md): L2 read traffic is not the limiter (the shared-memory port is: TMA writes +
    // packed-weight reads + UMMA B-operand reads ~ 48 KB per K-block), and lock-stepping the cluster costs more than
    // the saved L2 reads.  Opt-in for experiments: OB_GEMM_MC = 2 / 4.
    int mc = env.mc;
    if (mc < 1 || mc > 4 || (mc & (mc - 1)) || p.n_tiles % mc || BN % (8 * mc) || a.force_ctas > 0) mc = 1;
    // CTA-pair MMA (tcgen05 cta_group::2): 256 weight rows x 128 tokens per pair, each CTA stages half of the tokens.
    // OB_GEMM_2CTA = 0 / 1 overrides the automatic choice (large-M data-parallel problems).
    two = 0;  // opt-in until it beats the single-CTA kernel (profiles/r1_summary.md)
    if (env.has_two) two = (env.two != 0 && BN == 128 && p.n_tiles % 2 == 0 && a.force_ctas <= 0) ? 1 : 0;
    if (p.tail) { two = 0; mc = 1; }   // the fused tail is only wired into the plain single-CTA schedule
    if (two) mc = 2;
    p.mc = mc;
    if (mc > 1) cluster = (unsigned)mc;
  }
  if (p.mc < 1) p.mc = 1;
  CUtensorMap map;
  if (int e = make_act_map(&map, a.in_feats, a.M, a.K, BN)) return e;
  CUtensorMap mcmap = map;
  if (p.mc > 1) {
    if (int e = make_act_map(&mcmap, a.in_feats, a.M, a.K, BN / p.mc)) return e;
  }
  CUtensorMap wmap;
  p.w_rows2k = env.w2k;
  if (int e = make_w_map(&wmap, a.qweight, a.N, a.K, p.w_rows2k != 0)) return e;
#define OB_LAUNCH(bn)                                                              \
  case bn:                                                                         \
    return per_group ? launch<bn, true>(map, wmap, mcmap, p, grid, cluster, st) : launch<bn, false>(map, wmap, mcmap, p, grid, cluster, st);
  switch (BN) {
    OB_LAUNCH(16)
    OB_LAUNCH(32)
    OB_LAUNCH(64)
    case 128:
      if (two) return per_group ? launch<128, true, true>(map, wmap, mcmap, p, grid, cluster, st)
                                : launch<128, false, true>(map, wmap, mcmap, p, grid, cluster, st);
      return per_group ? launch<128, true>(map, wmap, mcmap, p, grid, cluster, st) : launch<128, false>(map, wmap, mcmap, p, grid, cluster, st);
    default:
      return OB_ERR_SHAPE;
  }
#undef OB_LAUNCH
}

}  // namespace ob
