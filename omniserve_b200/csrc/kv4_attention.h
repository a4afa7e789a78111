#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "w4a8_gemm.h"  // error codes

namespace ob {

struct KV4DecodeArgs {
  const __half* q; const __half* k; const __half* v;  // [B,Hq,128], [B,Hkv,128] views; head stride 128
  long long q_bs, k_bs, v_bs;                          // batch strides in elements
  __half* out;                                         // [B,Hq,128] contiguous
  const int64_t* retrieval_kv_pointers;                // [B,2,r_max_pages] absolute device addresses
  const int64_t* streaming_kv_pointers;                // [B,2,s_max_pages] or null
  int r_max_pages, s_max_pages;
  const int* lengths;                                  // [B] context incl. the new token, or null
  const int* retrieval_head_flags;                     // [Hkv] or null (= all retrieval)
  const int* head_rank_table;                          // [Hkv] or null (= identity)
  const int* dyn_idx; int dyn_pages;                   // [B,Hq,P] or null
  int B, Hq, Hkv, head_dim, tokens_per_block;
  int num_retrieval_kv_heads, num_streaming_kv_heads;
  int sink_tokens, local_tokens, sink_blocks, local_blocks;
  int timestep;                                        // max cached tokens in the batch
  int max_attended;                                    // upper bound of attended cached tokens per head
  int rotary_dim; float rotary_base; float rotary_scale;  // scale already inverted (1/factor)
  int force_split = 0;
  int8_t* q_out = nullptr; __half* q_scale = nullptr; __half* q_sum = nullptr;   // fused output quant (extension)
  int tokens_per_sub_chunk = 0;                        // > 0: fold the appended key into the page's kmax / kmin
  int hidden_dim_per_retrieval_token = 0;
  int stable_history = 0;                              // see AttnParams::stable_history
  const float* kv_scale_quant_orig = nullptr;          // non-null: per-tensor KV8 pages (device float[2]: K, V dequant scales)
  const float* kv_scale_orig_quant = nullptr;          //           and the quant scales used for the appended token
};

int kv4_decode_run(const KV4DecodeArgs& a, cudaStream_t st);

struct KV4PrefillArgs {
  __half* qkv;                                         // [T,(Hq+2Hkv)*128] in/out (RoPE in place on q,k)
  const int* seq_lens;                                 // [B]
  const int* padding_offset; int max_seq_len;          // [T]: padded index = t + padding_offset[t]
  const int64_t* retrieval_kv_pointers; const int64_t* streaming_kv_pointers;
  int r_max_pages, s_max_pages;
  const int* retrieval_head_flags; const int* head_rank_table;
  int T, B, Hq, Hkv;
  int num_retrieval_kv_heads, num_streaming_kv_heads;
  int sink_tokens, local_tokens, sink_blocks, local_blocks;
  int rotary_dim; float rotary_base; float rotary_scale;
  const float* kv_scale_orig_quant = nullptr;          // non-null: per-tensor KV8 pages (device float[2])
};
int kv4_prefill_write_run(const KV4PrefillArgs& a, cudaStream_t st);
// fused: the same page writes + the kmax / kmin statistics of every 16-token sub-chunk of the retrieval heads
int kv4_prefill_write_pool_run(const KV4PrefillArgs& a, int tokens_per_sub_chunk, cudaStream_t st);
int padding_offsets_run(int* out, const int* cu_seqlens, int B, int max_seq_len, cudaStream_t st);

}  // namespace ob
