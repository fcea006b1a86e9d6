// Internal (C++) launchers of the memory-bound kernels of KV-cached caption generation.
#pragma once
#include "common.cuh"

namespace b200 {

// cross K/V from the projection GEMM's [B*S, 2E] layout into head-major [B][H][S][hd] K and V
// planes (contiguous per (image, head): what the decode kernel streams).
int kv_to_head_major(const bf16* kv, bf16* k_hm, bf16* v_hm, int B, int S, int H, int hd, cudaStream_t s);
// single-query attention: for every group g (image) and head h, NQ query rows (beams) attend
// over the same nkeys keys.  q/o: [groups*nq, H*hd] rows; K/V: [groups][H][kv_len][hd], reading
// the first nkeys rows.  key_pad: optional [groups, nkeys] uint8.  pf_k / pf_v / pf_bytes: optional L2
// warm-up hint, the first pf_bytes of the K and V planes the NEXT call will stream; flags bit 0: load the
// K/V stream with the L2 evict-first priority, bit 1: fetch a short last chunk as 16-row boxes, bit 3: 3-deep
// rings in the fat-CTA shape (single query row per image only)
// fat_grid > 0: run as at most fat_grid one-per-SM CTAs of several producer/consumer units each, leaving the
// remaining SMs to concurrently running kernels; sched: optional two zeroed ints for dynamic (first come, first
// served) item distribution, left zeroed again by the launch (all five: tensor-core kernel only).
int attn_decode(const bf16* q, long long q_rs, const bf16* k, const bf16* v, int kv_len, int nkeys,
                bf16* o, long long o_rs, int groups, int nq, int H, int hd, const unsigned char* key_pad,
                float scale, cudaStream_t s, const bf16* pf_k = nullptr, const bf16* pf_v = nullptr,
                long long pf_bytes = 0, int flags = 0, int fat_grid = 0, int* sched = nullptr);
// self attention of one decode position with the cache append fused in: q/k/v of the current position
// are the three E-wide column blocks of qkv [R, 3E]; k/v are written to cache row `pos` of
// [R][H][max_len][hd] and attended together with rows [0,pos).
// seq (optional): [R, seq_ld] token ids of the hypotheses, columns [0, pos] valid; keys whose token == pad_idx are
// masked (the reference's tgt_key_padding_mask, decoder.py:162, rebuilt from the prefix on every generate() step).
int attn_decode_append(const bf16* qkv, long long qkv_rs, bf16* kcache, bf16* vcache, int max_len, int pos, bf16* o,
                       long long o_rs, int R, int H, int hd, float scale, cudaStream_t s,
                       const int64_t* seq = nullptr, long long seq_ld = 0, long long pad_idx = 0);
// p[r * stride + col] = src[r]
int store_col_i64(int64_t* p, const int64_t* src, long long n, long long stride, long long col, cudaStream_t s);
// greedy bookkeeping after a step: finished rows emit pad_id, END marks a row finished
int greedy_update(const int64_t* next_ids, int64_t* cur_tokens, int64_t* out_tokens, int* out_len,
                  unsigned char* finished, int* n_finished, int R, int max_len, int pos, long long end_id,
                  long long pad_id, cudaStream_t s);
// beam search: per image top-`beam` of (beam x V) candidate scores = score[beam] + log_softmax(logits)
int beam_topk(const float* logits, const float* beam_scores, const unsigned char* finished, int B, int beam,
              int V, long long end_id, int first_step, int64_t* out_tokens, int* out_parent,
              float* out_scores, cudaStream_t s);
int beam_advance(const int64_t* seq_in, int64_t* seq_out, const unsigned char* fin_in, unsigned char* fin_out,
                 const int64_t* tokens, const int* parent, int R, int beam, int max_len, int pos, long long end_id,
                 cudaStream_t s);
int beam_finalize(const int64_t* seqs, const float* scores, int B, int beam, int max_len, int n_tok,
                  long long end_id, long long pad_id, int64_t* out_tokens, int* out_len, float* out_score,
                  cudaStream_t s);
int fill_i64(int64_t* p, long long n, long long v, cudaStream_t s);
int fill_col_i64(int64_t* p, long long n, long long stride, long long v, cudaStream_t s);
// cache[dst row] = cache[parent row] for positions [0,pos]
int cache_reorder(const bf16* src_k, const bf16* src_v, bf16* dst_k, bf16* dst_v, const int* parent, int B,
                  int beam, int H, int hd, int max_len, int pos, cudaStream_t s);

}  // namespace b200
