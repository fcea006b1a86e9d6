// Internal (C++) launchers of the memory-bound kernels: embedding, LayerNorm, reductions,
// softmax-CE finalisation, optimizer.  All take raw device pointers and a stream.
#pragma once
#include "common.cuh"

namespace b200 {

// position of row r is (r % T) + t0  (t0 > 0: single-position decode steps)
// pos (optional): position of every row (packed / var-len batches)
int embed_pe_fwd(const int64_t* tokens, const float* emb, const float* pe, bf16* x, int B, int T,
                 int E, int V, float scale, cudaStream_t s, int t0 = 0, DropCfg dc = DropCfg{nullptr, 0u, 0u, 1.f},
                 const int32_t* pos = nullptr);
// packed (var-len) batches: sample b keeps its first cu[b+1]-cu[b] positions as rows [cu[b], cu[b+1]);
// ptok / ptgt / ppos = token id, target id and position of every kept row
int pack_rows(const int64_t* tokens, const int64_t* targets, const int32_t* cu, int B, int T, int64_t* ptok,
              int64_t* ptgt, int32_t* ppos, cudaStream_t s);
int embed_bwd(const int64_t* tokens, const bf16* dx, float* demb, int B, int T, int E, int V,
              long long pad_idx, float scale, cudaStream_t s, DropCfg dc = DropCfg{nullptr, 0u, 0u, 1.f});
// advances the dropout step counter (state[1]) on the stream
int drop_advance(uint32_t* state, cudaStream_t s);
int layernorm_fwd(const bf16* x, const float* gamma, const float* beta, bf16* y, float* mean,
                  float* rstd, int rows, int E, float eps, cudaStream_t s);
// y = LayerNorm(sum_s parts[s] + bias + residual): parts = fp32 split-K slabs of a "partials" GEMM
// (slab s at parts + s*slab_stride, row pitch ldp)
int layernorm_reduce_fwd(const float* parts, int nsplit, long long slab_stride, long long ldp, const float* bias,
                         const bf16* residual, long long ldr, const float* gamma, const float* beta, bf16* y,
                         int rows, int E, float eps, cudaStream_t s);
// dx_drop / dc (optional): second output dx * dropout-mask / (1 - p) of site dc (and dxsum sums THAT);
// dxsum (optional): += column sums of dx, i.e. the bias gradient of the Linear feeding this LayerNorm
int layernorm_bwd(const bf16* dy, const bf16* x, const float* gamma, const float* mean,
                  const float* rstd, bf16* dx, float* dgamma, float* dbeta, float* dxsum, int rows, int E,
                  cudaStream_t s, bf16* dx_drop = nullptr, DropCfg dc = DropCfg{nullptr, 0u, 0u, 1.f});
int colsum(const bf16* x, long long ldx, float* out, int M, int N, cudaStream_t s);
int cast_f32_to_bf16(const float* src, bf16* dst, long long n, cudaStream_t s);
int cast_bf16_to_f32(const bf16* src, float* dst, long long n, cudaStream_t s);

// softmax-CE finalisation of the LM-head partials written by the EPI_CE_FWD GEMM epilogue
int ce_finalize(const float* part_max, const float* part_sum, const float* tgt_logit,
                const int64_t* targets, int M, int n_tiles, long long ignore_index, float* row_lse,
                float* row_loss, float* loss_sum, float* valid_count, cudaStream_t s);
// out[0] = loss_sum/valid_count, out[1] = valid_count, out[2] = 1/valid_count
int ce_mean(const float* loss_sum, const float* valid_count, float* out, cudaStream_t s);
int argmax_finalize(const float* part_max, const float* part_idx, int M, int n_tiles,
                    int64_t* out_ids, float* out_max, cudaStream_t s);
// dlogits (fp32 [M,V]) -> bf16 copy (autograd compatibility path)
int grad_sumsq(const float* g, long long n, float* sumsq, cudaStream_t s);
// step_dev / lr_dev (optional): device-resident step counter (incremented here) and learning rate,
// so that a captured CUDA graph of the train step stays valid across replays
int adamw_step(float* p, bf16* p16, const float* g, float* m, float* v, long long n,
               const float* sumsq, float max_norm, float lr, float b1, float b2, float eps,
               float wd, int step, cudaStream_t s, int* step_dev = nullptr, const float* lr_dev = nullptr);

}  // namespace b200
