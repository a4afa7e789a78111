// W8A8 GEMM for sm_100a: C[M,N] (fp16) = (X[M,K] int8 . W[N,K]^T int8) * wscale[n] * ascale[m].
//
// Replaces /root/reference/kernels/csrc/qgemm/w8a8/w8a8_gemm_cuda.cu:537-600 (mma.sync m16n8k32 + cp.async, epilogue
// :515-530: `psum *= wscale * ascale` in fp32, one fp16 rounding) -- the GEMM of LServe's released W8A8KV8 setting
// (scripts/lserve_benchmark/launch.sh:6-7).  Same design as w4a8_gemm.cu minus the unpack stage: both operands are
// already INT8, so weights and activations are TMA-staged into 128B-swizzled shared memory and fed to
// tcgen05.mma kind::i8 as shared-memory descriptors (SS form), INT32 accumulators double-buffered in tensor memory.
// Orientation C^T = W . X^T: 128 weight rows = UMMA M (TMEM lanes), BN tokens = UMMA N.
// Warps: w0 TMA producer, w1 MMA issuer (+ TMEM allocation), w2-5 epilogue (one TMEM lane quarter each).
// Persistent over output tiles, token tiles of one weight tile adjacent in the schedule (weights stay in L2).
#include "launch.h"
#include "ptx.cuh"
#include "w4a8_gemm.h"

#include <algorithm>

namespace ob {
namespace w8 {

constexpr int BM = 128, BK = 128, NUM_THREADS = 192;
constexpr int A_STAGE = BM * BK;   // 16 KB

template <int BN>
struct Cfg {
  static constexpr int STAGES = BN >= 128 ? 5 : 6;
  static constexpr int B_STAGE = BN * BK;
  static constexpr int OUT_PITCH = BM + 8;                       // halves
  static constexpr int STAGING = BN * OUT_PITCH * 2;
  static constexpr int SMEM_A = 0;
  static constexpr int SMEM_B = SMEM_A + STAGES * A_STAGE;
  static constexpr int SMEM_STAGING = SMEM_B + STAGES * B_STAGE;
  static constexpr int SMEM_TOK = SMEM_STAGING + STAGING;        // float as[BN]
  static constexpr int SMEM_BAR = SMEM_TOK + BN * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int SMEM_MISC = SMEM_BAR + NUM_BARS * 8;
  static constexpr int SMEM_TOTAL = SMEM_MISC + 64 + 1024;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : 256));
  static_assert(SMEM_TOTAL <= 227 * 1024, "smem budget");
};

struct Params {
  const __half* wscales; const __half* ascales; __half* out;
  int M, N, K, ldc, n_tiles, m_tiles, kb;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
w8a8_gemm_kernel(const __grid_constant__ CUtensorMap act_map, const __grid_constant__ CUtensorMap w_map, const Params p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem + C::SMEM_A;
  uint8_t* sB = smem + C::SMEM_B;
  __half* stage16 = reinterpret_cast<__half*>(smem + C::SMEM_STAGING);
  float* sTok = reinterpret_cast<float*>(smem + C::SMEM_TOK);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::SMEM_BAR);
  uint64_t* full = bars;                      // both operand tiles of the stage landed
  uint64_t* empty = full + C::STAGES;         // MMAs that read the stage retired
  uint64_t* acc_full = empty + C::STAGES;     // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2], 4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::SMEM_MISC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&act_map); tma_prefetch_desc(&w_map);
    for (int i = 0; i < C::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles = p.n_tiles * p.m_tiles;

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();   // activations are the previous kernel's output (weights are not, but one producer keeps it simple)
      int st = 0, ph = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
        for (int kb = 0; kb < p.kb; ++kb) {
          mbar_wait(&empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&full[st], A_STAGE + C::B_STAGE);
          tma_load_2d(sA + st * A_STAGE, &w_map, kb * BK, nt * BM, &full[st]);
          tma_load_2d(sB + st * C::B_STAGE, &act_map, kb * BK, mt * BN, &full[st]);
          if (++st == C::STAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    int st = 0, ph = 0, acc = 0, acc_phase = 0;
    constexpr uint32_t idesc = umma_idesc_i8(BM, BN, true, true);
    const uint64_t adesc0 = umma_desc_kmajor_sw128(smem_u32(sA));
    const uint64_t bdesc0 = umma_desc_kmajor_sw128(smem_u32(sB));
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < p.kb; ++kb) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + (uint64_t)((st * A_STAGE) >> 4);
          const uint64_t bd = bdesc0 + (uint64_t)((st * C::B_STAGE) >> 4);
          umma_i8_ss(d_tmem, ad, bd, idesc, kb > 0 ? 1u : 0u);
          umma_i8_ss(d_tmem, ad + 2, bd + 2, idesc, 1u);
          umma_i8_ss(d_tmem, ad + 4, bd + 4, idesc, 1u);
          umma_i8_ss(d_tmem, ad + 6, bd + 6, idesc, 1u);
          umma_commit(&empty[st]);
          if (kb == p.kb - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++st == C::STAGES) { st = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================================================================ epilogue (warps 2-5 -> TMEM lane quarters 2,3,0,1)
    const int q = warp & 3;
    const int et = threadIdx.x - 64;   // 0..127
    int acc = 0, acc_phase = 0;
    pdl_wait();
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
      const int m0 = mt * BN;
      const int n_row = nt * BM + q * 32 + lane;
      const float wsc = n_row < p.N ? __half2float(p.wscales[n_row]) : 0.f;
      if (et < BN) sTok[et] = (m0 + et < p.M) ? __half2float(p.ascales[m0 + et]) : 0.f;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_acc + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float o = __int2float_rn((int)r[j]) * (wsc * sTok[c0 + j]);   // w8a8_gemm_cuda.cu:527-529
          stage16[(c0 + j) * C::OUT_PITCH + q * 32 + lane] = __float2half_rn(o);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
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
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int BN>
static int launch(const CUtensorMap& amap, const CUtensorMap& wmap, const Params& p, int grid, cudaStream_t st) {
  using C = Cfg<BN>;
  auto kern = w8a8_gemm_kernel<BN>;
  static bool attr_done[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_TOTAL) != cudaSuccess) return OB_ERR_CUDA;
    attr_done[dev] = true;
  }
  return launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), (size_t)C::SMEM_TOTAL, st, amap, wmap, p) == cudaSuccess ? 0 : OB_ERR_CUDA;
}

}  // namespace w8

int w8a8_gemm_run(const int8_t* in_feats, const int8_t* weight, const __half* wscales, const __half* ascales, __half* out,
                  int M, int N, int K, int ldc, cudaStream_t st) {
  using namespace w8;
  if (M <= 0) return 0;
  if (N % 8 != 0 || K % 128 != 0 || ldc % 8 != 0 || ldc < N) return OB_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(in_feats) & 15) || (reinterpret_cast<uintptr_t>(weight) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return OB_ERR_ALIGN;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return OB_ERR_ARG;
  const int sms = dev_sms(dev);
  const int BN = M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : 128;
  Params p{};
  p.wscales = wscales; p.ascales = ascales; p.out = out;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc;
  p.n_tiles = (N + BM - 1) / BM;
  p.m_tiles = (M + BN - 1) / BN;
  p.kb = K / BK;
  const int grid = std::min(p.n_tiles * p.m_tiles, sms);
  CUtensorMap amap, wmap;
  if (int e = make_act_map(&amap, in_feats, M, K, BN)) return e;
  if (int e = make_act_map(&wmap, weight, N, K, BM)) return e;    // [N, K] int8 row-major, box {128 B, 128 rows}, SW128
  switch (BN) {
    case 16: return launch<16>(amap, wmap, p, grid, st);
    case 32: return launch<32>(amap, wmap, p, grid, st);
    case 64: return launch<64>(amap, wmap, p, grid, st);
    default: return launch<128>(amap, wmap, p, grid, st);
  }
}

}  // namespace ob
