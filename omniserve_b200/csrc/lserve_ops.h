#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "w4a8_gemm.h"  // error codes

namespace ob {

struct PoolArgs {
  const __half* keys;                  // [T, H_in, Dh] post-RoPE keys (fp16)
  long long row_stride, head_stride;   // elements
  const int64_t* retrieval_kv_pointers;  // [B, 2, r_max_pages]
  int r_max_pages;
  const int* cu_seqlens;               // [B+1]
  const int* pooling_heads_idx;        // [num_pooling_heads] -> input head index
  int batch, num_pooling_heads, head_dim;
  int max_seqlen, pooling_size, page_size, size_per_retrieval_token;
  int kv_cache_with_zeros;
};
int paged_min_max_pool_run(const PoolArgs& a, cudaStream_t st);

struct SelectorArgs {
  const __half* q; long long q_bs;     // [B, Hq, Dh] view (head stride Dh), batch stride in elements
  __half* out;                         // [B, Hq, padded_sub_chunks(timestep)] -- zeroed by the op
  const int64_t* retrieval_kv_pointers; int r_max_pages;
  const int* lengths;                  // [B] incl. the new token, or null
  const int* retrieval_head_flags; const int* head_rank_table;
  int B, Hq, Hkv, head_dim, tokens_per_block;
  int size_per_retrieval_token, num_retrieval_kv_heads;
  int timestep;
  int rotary_dim; float rotary_base; float rotary_scale;  // scale already inverted (1/factor)
  int tokens_per_sub_chunk, hidden_dim_per_retrieval_token;
};
int page_selector_run(const SelectorArgs& a, cudaStream_t st);

// Device-side page choice after the selector (decoding_attention.py:132-141): scores fp16 [rows, pitch] sub-chunk scores,
// out int32 [rows, k_out] = the k_out-1 best pages among pages 0..total_pages-2 (page score = max of its sub-chunks) followed
// by the newest page total_pages-1.
int page_topk_run(const __half* scores, int* out, int rows, int pitch_sub_chunks, int sub_chunks_per_page, int total_pages,
                  int k_out, cudaStream_t st);

}  // namespace ob
