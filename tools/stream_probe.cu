// Stand-alone weight-streaming probe (sm_100a): what bounds the per-CTA K-block rate of the decode GEMM's load
// pipeline, and how much of a launch's fixed cost does PDL + co-residency hide?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_build/stream_probe tools/stream_probe.cu
// Every CTA streams KB "K-blocks" of 8 KB of packed W4 weights (the reference layout [N/32][K/32][512 B], tile =
// 128 rows x 128 K) into a shared-memory ring; a consumer warp releases each stage as soon as it has landed (no compute).
//   load = 0: one 3-D TMA per stage (16 rows of 512 B)      [what w4a8_gemm.cu does]
//   load = 1: one 3-D TMA per stage (4 rows of 2 KB)
//   load = 2: four cp.async.bulk of 2 KB per stage
//   load = 3: one cp.async.bulk of 8 KB per stage (re-tiled layout: a tile's K-blocks contiguous)
//   act  = 1: plus one 2-D TMA of an activation tile (BN rows x 128 B, 128B swizzle) per K-block in its own ring
// Chain mode: N back-to-back launches with / without programmatic dependent launch, consumer gated by
// griddepcontrol.wait, plus a fixed "epilogue" spin, to see what co-residency hides.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../omniserve_b200/csrc/launch.h"
#include "../omniserve_b200/csrc/ptx.cuh"

using namespace ob;

struct P {
  const uint8_t* w;
  int kb_per_cta, k32_total;   // K/32 of the whole matrix
  int stages, a_stages, load, act, bn;
  int gate;                    // chain mode: consumer waits for the grid dependency before consuming
  int tail_cycles;             // chain mode: spin after the last K-block (epilogue stand-in)
  int consumers;               // consumer warps (each must arrive on the empty barrier)
};

__global__ void __launch_bounds__(256, 1)
probe_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap wmap2k,
             const __grid_constant__ CUtensorMap amap, const P p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + p.stages * 8192;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (p.act ? p.a_stages * p.bn * 128 : 0));
  uint64_t* w_full = bars;
  uint64_t* w_empty = w_full + p.stages;
  uint64_t* a_full = w_empty + p.stages;
  uint64_t* a_empty = a_full + p.a_stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], p.consumers); }
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    mbar_fence_init();
  }
  __syncthreads();
  const int nt = blockIdx.x;
  if (warp == 0) {
    if (lane == 0) {
      int st = 0, ph = 0;
      for (int kb = 0; kb < p.kb_per_cta; ++kb) {
        mbar_wait(&w_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&w_full[st], 8192);
        uint8_t* dst = sW + st * 8192;
        if (p.load == 0) tma_load_3d(dst, &wmap, 0, kb * 4, nt * 4, &w_full[st]);
        else if (p.load == 1) tma_load_3d(dst, &wmap2k, 0, kb, nt * 4, &w_full[st]);
        else if (p.load == 2) {
          for (int r = 0; r < 4; ++r)
            bulk_g2s(dst + r * 2048, p.w + ((size_t)(nt * 4 + r) * p.k32_total + kb * 4) * 512, 2048, &w_full[st]);
        } else {
          bulk_g2s(dst, p.w + ((size_t)nt * p.kb_per_cta + kb) * 8192, 8192, &w_full[st]);
        }
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && p.act) {
      if (p.gate) pdl_wait();
      int st = 0, ph = 0;
      for (int kb = 0; kb < p.kb_per_cta; ++kb) {
        mbar_wait(&a_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&a_full[st], p.bn * 128);
        tma_load_2d(sA + st * p.bn * 128, &amap, kb * 128, 0, &a_full[st]);
        if (++st == p.a_stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 2 && warp < 2 + p.consumers) {
    if (p.gate) pdl_wait();
    int st = 0, ph = 0, ast = 0, aph = 0;
    uint32_t sink = 0;
    for (int kb = 0; kb < p.kb_per_cta; ++kb) {
      mbar_wait(&w_full[st], ph);
      sink += lds_u32(smem_u32(sW + st * 8192 + lane * 16 + (warp - 2) * 2048));
      if (p.act && warp == 2) {
        mbar_wait(&a_full[ast], aph);
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_empty[ast]);
        if (++ast == p.a_stages) { ast = 0; aph ^= 1; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&w_empty[st]);
      if (++st == p.stages) { st = 0; ph ^= 1; }
    }
    if (p.tail_cycles > 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < p.tail_cycles) {}
    }
    if (sink == 0x12345678u) printf("x");
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled enc() {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  return (PFN_encodeTiled)f;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct Run {
  int grid, kb, stages, a_stages, load, act, bn, consumers, pad_smem, chain, pdl, gate, tail;
};

static double run(const Run& r, uint8_t** wbufs, int nbuf, uint8_t* act, size_t wbytes_alloc) {
  const int n_tiles = r.grid;
  const int K = r.kb * 128, N = n_tiles * 128;
  const size_t bytes = (size_t)N * K / 2;
  if (bytes > wbytes_alloc) { printf("skip (too large)\n"); return 0; }
  PFN_encodeTiled e = enc();
  std::vector<CUtensorMap> wm(nbuf), wm2(nbuf);
  for (int i = 0; i < nbuf; ++i) {
    cuuint64_t dims[3] = {64, (cuuint64_t)(K / 32), (cuuint64_t)(N / 32)};
    cuuint64_t str[2] = {512, (cuuint64_t)(K / 32) * 512};
    cuuint32_t box[3] = {64, 4, 4}, es[3] = {1, 1, 1};
    if (e(&wm[i], CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, wbufs[i], dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc fail\n"); exit(1); }
    cuuint64_t d2[3] = {256, (cuuint64_t)(K / 128), (cuuint64_t)(N / 32)};
    cuuint64_t s2[2] = {2048, (cuuint64_t)(K / 32) * 512};
    cuuint32_t b2[3] = {256, 1, 4};
    if (e(&wm2[i], CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, wbufs[i], d2, s2, b2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc fail\n"); exit(1); }
  }
  CUtensorMap am;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)r.bn};
    cuuint64_t str[1] = {(cuuint64_t)K};
    cuuint32_t box[2] = {128, (cuuint32_t)r.bn}, es[2] = {1, 1};
    if (e(&am, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, act, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc fail\n"); exit(1); }
  }
  size_t smem = (size_t)r.stages * 8192 + (r.act ? (size_t)r.a_stages * r.bn * 128 : 0) + 16 * (r.stages + r.a_stages) + 2048;
  if ((int)smem < r.pad_smem) smem = r.pad_smem;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  auto launch_one = [&](int i) {
    P p{};
    p.w = wbufs[i % nbuf]; p.kb_per_cta = r.kb; p.k32_total = K / 32; p.stages = r.stages; p.a_stages = r.a_stages;
    p.load = r.load; p.act = r.act; p.bn = r.bn; p.gate = r.gate; p.tail_cycles = r.tail; p.consumers = r.consumers;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(r.grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = r.pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, probe_kernel, wm[i % nbuf], wm2[i % nbuf], am, p));
  };
  const int per_graph = r.chain > 0 ? r.chain : nbuf;
  for (int i = 0; i < per_graph; ++i) launch_one(i);
  CK(cudaStreamSynchronize(st));
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
  for (int i = 0; i < per_graph; ++i) launch_one(i);
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(a, st));
    CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(b, st));
    CK(cudaStreamSynchronize(st));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(st);
  const double us = best * 1e3 / per_graph;
  printf("grid=%4d kb=%4d stages=%2d load=%d act=%d(bn=%3d,as=%d) cons=%d smem=%6zu pdl=%d gate=%d tail=%5d chain=%2d : %8.2f us/launch  %7.1f GB/s  %6.3f us/kb\n",
         r.grid, r.kb, r.stages, r.load, r.act, r.bn, r.a_stages, r.consumers, smem, r.pdl, r.gate, r.tail, r.chain, us,
         (double)bytes / us / 1e3, us / r.kb);
  fflush(stdout);
  return us;
}

int main(int argc, char** argv) {
  const size_t WB = 160ull << 20;
  const int NBUF = 4;
  uint8_t* wb[NBUF];
  for (int i = 0; i < NBUF; ++i) { CK(cudaMalloc(&wb[i], WB)); CK(cudaMemset(wb[i], i + 1, WB)); }
  uint8_t* act;
  CK(cudaMalloc(&act, 128 * 16384)); CK(cudaMemset(act, 1, 128 * 16384));
  const char* what = argc > 1 ? argv[1] : "all";
  auto is = [&](const char* s) { return !strcmp(what, "all") || !strcmp(what, s); };
  //        grid  kb  st as ld act bn cons pad chain pdl gate tail
  if (is("rate")) {
    printf("# per-CTA K-block rate of the pure load pipeline (112 K-blocks per CTA)\n");
    for (int grid : {32, 148}) for (int load : {0, 1, 2, 3}) for (int stages : {4, 8, 16, 24})
      run({grid, 112, stages, 8, load, 0, 64, 1, 0, 0, 1, 0, 0}, wb, NBUF, act, WB);
    printf("# with the activation ring (BN = 64 / 16)\n");
    for (int grid : {32, 148}) for (int bn : {16, 64}) for (int as : {4, 8})
      run({grid, 112, 16, as, 0, 1, bn, 1, 0, 0, 1, 0, 0}, wb, NBUF, act, WB);
    printf("# 4 consumer warps\n");
    for (int load : {0, 3}) run({148, 112, 16, 8, load, 0, 64, 4, 0, 0, 1, 0, 0}, wb, NBUF, act, WB);
    printf("# two CTAs per SM (8 / 12 stages each)\n");
    for (int load : {0, 2, 3}) for (int stages : {6, 8, 12})
      run({296, 56, stages, 4, load, 0, 64, 1, 0, 0, 1, 0, 0}, wb, NBUF, act, WB);
    for (int stages : {6, 8}) run({296, 56, stages, 4, 0, 1, 64, 1, 0, 0, 1, 0, 0}, wb, NBUF, act, WB);
    printf("# three / four CTAs per SM\n");
    run({444, 37, 6, 4, 0, 0, 64, 1, 0, 0, 1, 0, 0}, wb, NBUF, act, WB);
    run({592, 28, 6, 4, 0, 0, 64, 1, 0, 0, 1, 0, 0}, wb, NBUF, act, WB);
    run({592, 28, 6, 4, 3, 0, 64, 1, 0, 0, 1, 0, 0}, wb, NBUF, act, WB);
  }
  if (is("chain")) {
    printf("# chain of 8 launches, 8 K-blocks per CTA (9.7 MB per launch), consumer gated by griddepcontrol.wait\n");
    for (int kb : {8, 32}) for (int tail : {0, 4000}) {
      run({148, kb, 8, 4, 0, 0, 64, 1, 0, 8, 0, 0, tail}, wb, NBUF, act, WB);            // no PDL
      run({148, kb, 8, 4, 0, 0, 64, 1, 0, 8, 1, 1, tail}, wb, NBUF, act, WB);            // PDL, co-resident (72 KB)
      run({148, kb, 8, 4, 0, 0, 64, 1, 120000, 8, 1, 1, tail}, wb, NBUF, act, WB);       // PDL, no room for a 2nd CTA
    }
  }
  return 0;
}
