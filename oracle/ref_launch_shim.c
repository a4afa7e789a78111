/* TEST INFRASTRUCTURE (oracle/): LD_PRELOAD interposer used ONLY when the reference's own page-selector kernel
 * (oracle/_ref/omniserve_backend/fused_attention_selector.so, built unmodified from /root/reference) is run as a
 * second opinion for SURVEY.md section 8 row a9.
 *
 * Why: the reference launches that kernel with ZERO bytes of dynamic shared memory
 * (sparse_utils/KVPageSelector/KVPageSelectorTemplate.hpp:1345-1347 `smem_size_in_bytes` returns 0) while the kernel
 * stages the rotated query through `extern __shared__ char smem_[]` (`q_smem_`, :1001-1053: rotary_dim halves = 256 B).
 * Those stores land beyond the CTA's shared-memory window; on sm_100 the window is enforced and the launch dies with
 * "illegal memory access" (observed on B200 in round 1; compute-sanitizer in round 2: "Invalid __shared__ write of size
 * 8 bytes ... Access at 0x580 is out of bounds", 0x580 = the kernel's 1408 B of static shared memory,
 * profiles/r2_selector_memcheck.log).  The reference
 * sources stay untouched: this shim adds OB_REF_EXTRA_SMEM bytes (default 1024) of dynamic shared memory to kernel
 * launches that ask for none, which is what a correct `smem_size_in_bytes` would have requested.
 *
 *   gcc -shared -fPIC -O2 -o oracle/_ref/libref_launch_shim.so oracle/ref_launch_shim.c -ldl
 *   LD_PRELOAD=oracle/_ref/libref_launch_shim.so python tests/ref_selector_worker.py in.npz out.npz
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stddef.h>
#include <stdlib.h>

typedef struct { unsigned int x, y, z; } dim3_t;
typedef int (*launch_fn)(const void*, dim3_t, dim3_t, void**, size_t, void*);

int cudaLaunchKernel(const void* func, dim3_t grid, dim3_t block, void** args, size_t smem, void* stream) {
  static launch_fn real = 0;
  static size_t extra = 0;
  if (!real) {
    real = (launch_fn)dlsym(RTLD_NEXT, "cudaLaunchKernel");
    const char* e = getenv("OB_REF_EXTRA_SMEM");
    extra = e ? (size_t)atol(e) : 1024;
  }
  if (smem == 0) smem = extra;
  return real(func, grid, block, args, smem, stream);
}
