// Programmatic dependent launch (PDL) helpers.  Every kernel of this library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization so that, inside the decode step's stream / CUDA graph,
// kernel N+1 may become resident while kernel N drains.  Device side: pdl_trigger() at kernel entry lets the
// next kernel in; pdl_wait() must precede the first access to anything the previous kernel produced.
#pragma once
#include <cuda_runtime.h>

namespace ob {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace ob
