/*
 * b200_decoder.h — C ABI of the B200-native caption-decoder hot path.
 *
 * The reference (wazzuck/multimodal-image-transformer) has no FFI of its own: its hot path is the
 * Python surface decoder.TransformerDecoder.forward (decoder.py:134-193), the train step
 * (train.py:71-120) and ImageToTextModel.generate (model.py:171-255), all of which bottom out in
 * torch.nn library calls.  This header is the boundary a maintainer would bind (ctypes, see
 * INTEGRATION.md) to replace those library calls with hand-written sm_100a kernels.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - every entry point returns 0 on success, a negative code on failure and never throws;
 *     b200_last_error() returns a description of the last failure on the calling thread;
 *   - no allocation inside the library except the engine's own workspace, which is sized by
 *     b200_engine_workspace_bytes() and supplied by the caller;
 *   - bf16 storage / fp32 accumulate.  Matrices are row-major with an explicit leading
 *     dimension counted in elements.
 */
#ifndef B200_DECODER_H
#define B200_DECODER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

/* ---- library ------------------------------------------------------------------------------ */
int b200_version(void);
const char* b200_last_error(void);
/* 0 when device `dev` is an sm_100 part and the kernels can run on it. */
int b200_check_device(int dev);
/* number of kernels this library has launched in the calling process (bench.py "gpu_launches"). */
long long b200_launch_count(void);
/* per-launch device timing of the tcgen05 GEMM family (bench.py "roofline"): between begin and
 * end every GEMM launch is bracketed by two CUDA events on its own stream; end (after the caller
 * has synchronised) returns the launch count, summed milliseconds and summed algorithmic flops
 * (2*M*N*K per launch), optionally per launch. */
int b200_gemm_profile_begin(int32_t max_launches);
int b200_gemm_profile_end(int32_t* n_launches, double* total_ms, double* total_flops,
                          float* per_launch_ms, double* per_launch_flops, int32_t cap);

/* ---- GEMM: D[M,N] = epi(A[M,K] . B[N,K]^T)  (replaces aten::addmm behind nn.Linear,
 *      torch/nn/functional.py:5798-5875 in-projections, decoder.py:191 fc_out and their
 *      autograd dgrad/wgrad).  tcgen05.mma + TMEM accumulators, TMA-fed, fused epilogue.
 *   a_mn_major = 0: A is [M,K] with K contiguous (lda = row pitch).
 *   a_mn_major = 1: A is stored transposed, [K,M] with M contiguous (lda = pitch of a K row).
 *   b_mn_major likewise for B ([N,K] K-contiguous vs [K,N] N-contiguous).
 *   epilogue order: acc (+bias[n]) -> act -> (* (relu_mask[m,n] > 0)) -> (+ residual[m,n]).
 *   d_fp32 = 0: D is bf16.  d_fp32 = 1: D is fp32; with accumulate = 1 the tile is added to D
 *   with red.global.add (required when split_k > 1; caller zero-fills first).
 */
enum { B200_ACT_NONE = 0, B200_ACT_RELU = 1, B200_ACT_GELU = 2 };

typedef struct {
  int32_t M, N, K;
  const void* A; int64_t lda; int32_t a_mn_major;
  const void* B; int64_t ldb; int32_t b_mn_major;
  void* D; int64_t ldd; int32_t d_fp32; int32_t accumulate;
  const float* bias;                        /* [N] fp32 or NULL */
  const void* residual; int64_t ldr;        /* bf16 [M,N] or NULL */
  const void* relu_mask; int64_t ldm;       /* bf16 [M,N] or NULL */
  int32_t act;
  int32_t split_k;                          /* >=1; 0 = choose (fp32 accumulate outputs only) */
  int32_t block_n;                          /* 0 = choose, else 128 or 256 */
} b200_gemm_args;
int b200_gemm(const b200_gemm_args* a, void* stream);

/* Plain CUDA-core GEMM with the same contract (no split_k): test instrument for the tcgen05
 * kernel at sizes where a CPU oracle would take too long.  Not used by the product path. */
int b200_gemm_check(const b200_gemm_args* a, void* stream);

/* ---- LM head fused with softmax cross-entropy (decoder.py:191 + train.py:90,327):
 *      logits are never written.  x [M,E] bf16, w [V,E] bf16, bias [V] fp32, targets [M] int64.
 *   fwd: row_lse[M], row_loss[M] (0 where target == ignore_index), loss_sum[1], valid_count[1].
 *        scratch: fp32, at least 3 * M * ceil(V/256) + M elements.
 *   bwd: dlogits[M,ldd] bf16 = (softmax - onehot) * (*inv_count) for valid rows, 0 otherwise. */
int b200_lmhead_ce_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                       const int64_t* targets, int32_t M, int32_t V, int32_t E,
                       int64_t ignore_index, float* row_lse, float* row_loss, float* loss_sum,
                       float* valid_count, float* scratch, void* stream);
int b200_lmhead_ce_bwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                       const int64_t* targets, int32_t M, int32_t V, int32_t E,
                       int64_t ignore_index, const float* row_lse, const float* inv_count,
                       void* dlogits, int64_t ldd, void* stream);
/* Greedy head: argmax_v(x.w^T + b) per row, first index on ties (torch.argmax, model.py:233).
 * scratch: fp32, at least 2 * M * ceil(V/128) elements. */
int b200_lmhead_argmax(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                       int32_t M, int32_t V, int32_t E, int64_t* out_ids, float* out_max,
                       float* scratch, void* stream);

/* ---- elementwise / normalisation kernels ---------------------------------------------------- */
/* x[b,t,:] = emb[tok[b,t],:] * scale + pe[t,:]   (decoder.py:168-170, 71-72); fp32 table. */
int b200_embed_pe_fwd(const int64_t* tokens, const float* emb, const float* pe, void* x_bf16,
                      int32_t B, int32_t T, int32_t E, int32_t V, float scale, void* stream);
/* demb[tok,:] += scale * dx[b,t,:]  for tok != pad_idx (nn.Embedding padding_idx, decoder.py:105) */
int b200_embed_bwd(const int64_t* tokens, const void* dx_bf16, float* demb, int32_t B, int32_t T,
                   int32_t E, int32_t V, int64_t pad_idx, float scale, void* stream);
/* y = LayerNorm(x) * gamma + beta, eps inside sqrt, biased variance (torch LayerNorm). */
int b200_layernorm_fwd(const void* x_bf16, const float* gamma, const float* beta, void* y_bf16,
                       float* mean, float* rstd, int32_t rows, int32_t E, float eps, void* stream);
/* dx (bf16), dgamma/dbeta (fp32, accumulated with atomics: caller zero-fills).  dxsum (optional,
 * fp32 [E], accumulated) receives the column sums of dx: the bias gradient of the Linear whose
 * output (plus residual) this LayerNorm normalised. */
int b200_layernorm_bwd(const void* dy_bf16, const void* x_bf16, const float* gamma,
                       const float* mean, const float* rstd, void* dx_bf16, float* dgamma,
                       float* dbeta, float* dxsum, int32_t rows, int32_t E, void* stream);
/* out[n] += sum_m x[m,n]   (bias gradients). x bf16 [M,N]. */
int b200_colsum(const void* x_bf16, int64_t ldx, float* out, int32_t M, int32_t N, void* stream);
int b200_cast_f32_to_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
int b200_cast_bf16_to_f32(const void* src_bf16, float* dst, int64_t n, void* stream);

/* ---- attention (torch/nn/functional.py:6244-6700 via nn.MultiheadAttention) ----------------
 * q/k/v/o are bf16 with explicit (batch, position) strides in elements; head h occupies columns
 * [h*hd, (h+1)*hd).  hd in {32, 64, 96, 128}.  lse [B,H,Tq] fp32 is saved for backward.
 * Self attention: causal (j <= i) + key padding (tokens[b,j] == pad_idx), decoder.py:158-162.
 * Cross attention: optional key padding mask mem_pad [B,S] (uint8, 1 = masked), model.py:158. */
typedef struct {
  const void* q; int64_t q_bs, q_ts;
  const void* k; int64_t k_bs, k_ts;
  const void* v; int64_t v_bs, v_ts;
  void* o; int64_t o_bs, o_ts;
  float* lse;
  int32_t B, H, Tq, Tk, hd;
  int32_t causal;
  const int64_t* key_tokens; int64_t pad_idx;   /* self: key padding from token ids (or NULL) */
  const uint8_t* key_pad_mask;                  /* cross: [B,Tk] or NULL */
  float scale;                                  /* 1/sqrt(hd) */
  /* Packed (var-len) sequences -- the reference pads every caption to MAX_SEQ_LEN (tokenizer.py:293-313,
   * dataset.py:176-206); packing drops the PAD rows.  cu_q / cu_k: device int32 [B+1] row offsets of every sample in
   * the q-side (q, o, dO, dq) / k-side (k, v, dk, dv) tensors, sample b owns rows [cu[b], cu[b+1]); total_q / total_k =
   * cu[B].  Tq / Tk are then the LARGEST per-sample lengths, lse stays [B,H,Tq], batch strides of packed tensors are
   * ignored.  NULL (the default) = regular [B, T] batches. */
  const int32_t* cu_q; const int32_t* cu_k;
  int32_t total_q, total_k;
} b200_attn_fwd_args;
int b200_attn_fwd(const b200_attn_fwd_args* a, void* stream);

typedef struct {
  b200_attn_fwd_args f;                         /* same tensors as forward (o, lse are inputs) */
  const void* d_o; int64_t do_bs, do_ts;
  void* dq; int64_t dq_bs, dq_ts;
  void* dk; int64_t dk_bs, dk_ts;
  void* dv; int64_t dv_bs, dv_ts;
} b200_attn_bwd_args;
int b200_attn_bwd(const b200_attn_bwd_args* a, void* stream);
/* Bring-up instrument of the tcgen05 attention kernels (head dim 64, Tq <= 64: the shapes b200_attn_fwd / _bwd route to
   csrc/attention_tc.cu): device buffer of 32 x 16 int64 that CTA 0 of the following launches fills with clock64 stamps of
   its pipeline events (producer issue, S issued, P seen, PV issued, softmax phases ...); NULL switches it off. */
int b200_attn_tc_trace(void* stamps_dev);

/* ---- optimizer (train.py:96-100, 319-325): global-norm clip + AdamW over one flat arena ---- */
/* sumsq[0] += sum g^2 ; caller zero-fills. */
int b200_grad_sumsq(const float* grad, int64_t n, float* sumsq, void* stream);
/* torch.optim.AdamW semantics; clip coefficient min(1, max_norm / (sqrt(*sumsq) + 1e-6)) is
 * applied to g on the fly (max_norm <= 0 disables).  Writes fp32 master and bf16 shadow.
 * step is the 1-based step count. */
int b200_adamw_step(float* param, void* param_bf16, const float* grad, float* exp_avg,
                    float* exp_avg_sq, int64_t n, const float* sumsq, float max_norm, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int32_t step,
                    void* stream);

/* Same update with the step counter (incremented by the call) and the learning rate read from
 * device memory: a CUDA graph that captured the train step stays valid across replays. */
int b200_adamw_step_dev(float* param, void* param_bf16, const float* grad, float* exp_avg,
                        float* exp_avg_sq, int64_t n, const float* sumsq, float max_norm,
                        const float* lr_dev, float beta1, float beta2, float eps, float weight_decay,
                        int32_t* step_dev, void* stream);

/* ---- decoder engine: whole-model entry points ----------------------------------------------
 * The engine owns no parameters: the caller binds flat arenas whose layout is fixed by
 * b200_engine_param_offset().  Tensor names are the reference state_dict keys relative to the
 * decoder ("token_embedding.weight", "transformer_decoder.layers.0.self_attn.in_proj_weight", …,
 * "fc_out.bias"), plus "projection.weight"/"projection.bias" when enc_dim != embed_dim. */
typedef struct {
  int32_t vocab_size, embed_dim, num_heads, num_layers, ff_dim, max_seq_len;
  int32_t enc_dim;          /* memory feature width; == embed_dim -> Identity projection */
  int64_t pad_idx;
  float ln_eps;
  int32_t act;              /* B200_ACT_RELU for parity with nn.TransformerDecoderLayer */
} b200_engine_config;

typedef struct b200_engine b200_engine;

int b200_engine_create(const b200_engine_config* cfg, b200_engine** out);
void b200_engine_destroy(b200_engine* e);
/* number of fp32 elements of the flat parameter arena / offset+numel of a named tensor (-1 if
 * the name is unknown).  Every tensor starts on a 64-element boundary. */
int64_t b200_engine_param_count(const b200_engine* e);
int64_t b200_engine_param_offset(const b200_engine* e, const char* name, int64_t* numel);
int b200_engine_param_name(const b200_engine* e, int32_t index, char* buf, int32_t buflen);
int32_t b200_engine_num_params(const b200_engine* e);
/* params_f32/grads_f32/params_bf16: arenas of b200_engine_param_count() elements. */
int b200_engine_bind(b200_engine* e, float* params_f32, void* params_bf16, float* grads_f32,
                     const float* pe_f32);
/* Element type of the `memory` argument of every later call: 0 = fp32 (what the reference's encoder
 * hands over, model.py:133-151; cast to bf16 by the engine), 1 = bf16 (pre-cast / cached frozen-encoder
 * features: consumed in place, so the buffer must stay valid and unchanged until the backward of the
 * same step has run; half the host->device bytes).  dmemory stays fp32. */
int b200_engine_set_memory_dtype(b200_engine* e, int32_t is_bf16);
/* Dropout of the training forward (reference config.py:69 DROPOUT = 0.1; modules: decoder.py:72,
 * torch/nn/modules/transformer.py:1175,1195,1199, attention-probability dropout functional.py:6682).
 * p = 0 disables it.  state_dev: caller-owned device buffer of two uint32 [seed, counter]; every
 * forward with training = 1 advances the counter once and derives all 1 + 6 L masks from
 * (seed, counter, site, element index) without storing them; backward regenerates the same masks.
 * Forwards with training = 0 and generation never drop. */
int b200_engine_set_dropout(b200_engine* e, float p, uint32_t* state_dev);
/* `mem_dim` is the width of the memory rows handed to the forward calls: embed_dim (already
 * projected, the decoder.TransformerDecoder.forward contract) or enc_dim (the engine applies
 * model.py:145's projection itself).  training = 1 keeps every activation backward needs. */
int64_t b200_engine_workspace_bytes(const b200_engine* e, int32_t B, int32_t T, int32_t S,
                                    int32_t mem_dim, int32_t training);
int b200_engine_set_workspace(b200_engine* e, void* ws, int64_t bytes);

/* decoder.TransformerDecoder.forward (decoder.py:134-193): tokens [B,T] int64, memory
 * [B,S,mem_dim] fp32, mem_pad [B,S] uint8 (1 = padded) or NULL -> logits [B,T,V] fp32. */
int b200_engine_forward_logits(b200_engine* e, const int64_t* tokens, const float* memory,
                               const uint8_t* mem_pad, int32_t B, int32_t T, int32_t S,
                               int32_t mem_dim, int32_t training, float* logits, void* stream);
/* forward + fused LM-head/CE (train.py:83-90, 142-145), logits never materialised:
 * loss_out[0] = mean CE over targets != ignore_index, loss_out[1] = number of such targets. */
int b200_engine_forward_loss(b200_engine* e, const int64_t* tokens, const int64_t* targets,
                             const float* memory, const uint8_t* mem_pad, int32_t B, int32_t T,
                             int32_t S, int32_t mem_dim, int64_t ignore_index, int32_t training,
                             float* loss_out, void* stream);
/* The same on a PACKED (var-len) batch.  The reference pads every caption to MAX_SEQ_LEN (tokenizer.py:293-313,
 * dataset.py:176-206), so 75-85 % of the positions of a real batch are PAD; here sample b keeps only its first
 * cu_seqlens[b+1] - cu_seqlens[b] positions (its non-PAD prefix) and every GEMM / LayerNorm / CE row-wise kernel runs on
 * total_rows = cu_seqlens[B] rows; the attention kernels index per-sample offsets.  tokens / targets stay [B,T] (the
 * engine gathers them), cu_seqlens is a device int32 [B+1] the caller keeps alive until backward is done.  Loss and
 * gradients equal the padded call's (PAD targets are ignored and PAD keys masked there).  The backward entry points
 * below run unchanged on the packed activations. */
int b200_engine_forward_loss_packed(b200_engine* e, const int64_t* tokens, const int64_t* targets,
                                    const float* memory, const uint8_t* mem_pad, int32_t B, int32_t T, int32_t S,
                                    int32_t mem_dim, int64_t ignore_index, int32_t training,
                                    const int32_t* cu_seqlens, int32_t total_rows, float* loss_out, void* stream);
/* backward of the last forward_loss(training=1) into the bound gradient arena (accumulating:
 * the caller zero-fills, as optimizer.zero_grad() does in train.py:80).
 * inv_count_dev: optional device scalar replacing 1/valid_count (data-parallel global mean).
 * dmemory: optional [B,S,embed_dim] fp32 gradient w.r.t. memory (mem_dim == embed_dim only).
 * bucket_events: optional array of cudaEvent_t, one per gradient bucket (see
 * b200_engine_grad_buckets), recorded on `stream` right after the bucket's last write. */
int b200_engine_backward(b200_engine* e, const float* inv_count_dev, float* dmemory,
                         void* const* bucket_events, int32_t num_bucket_events, void* stream);
/* the same backward cut into the L + 2 parts that match the gradient buckets (part 0 = LM head,
 * part k = layer L-k, part L+1 = embedding + projection): runs parts [first_part, last_part].
 * The data-parallel driver interleaves one all-reduce per part (capturable in a CUDA graph). */
int b200_engine_backward_parts(b200_engine* e, const float* inv_count_dev, float* dmemory,
                               int32_t first_part, int32_t last_part, void* stream);
/* backward of the last forward_logits(training=1) from an explicit dlogits [B,T,V] fp32: the
 * autograd-compatibility path behind decoder.TransformerDecoder.forward + loss.backward(). */
int b200_engine_backward_from_dlogits(b200_engine* e, const float* dlogits, float* dmemory,
                                      void* stream);
/* gradient buckets for the data-parallel all-reduce, in the order they become final during
 * backward: fc_out, layer L-1 .. layer 0, embedding(+projection).  Fills up to `cap`
 * (offset,count) pairs in arena elements; returns the number of buckets (L + 2). */
int32_t b200_engine_grad_buckets(const b200_engine* e, int64_t* offsets, int64_t* counts,
                                 int32_t cap);

/* ---- KV-cached generation (replaces the full-prefix recompute loop of model.py:216-242) ------
 * decode_begin projects the image memory ONCE per image into per-layer cross-attention K/V
 * (head-major, shared by all beams of an image) and carves the self-attention cache from `ws`
 * (b200_engine_decode_workspace_bytes, 256-byte aligned).  Rows are image-major: row = b*beam+k. */
int64_t b200_engine_decode_workspace_bytes(const b200_engine* e, int32_t B, int32_t beam, int32_t S,
                                           int32_t mem_dim, int32_t max_len);
int b200_engine_decode_begin(b200_engine* e, const float* memory, const uint8_t* mem_pad, int32_t B,
                             int32_t beam, int32_t S, int32_t mem_dim, int32_t max_len, void* ws,
                             int64_t ws_bytes, void* stream);
/* how the current decode plan is scheduled (after decode_begin): info[0] = concurrent image partitions,
 * info[1] = SMs budgeted to the cross-attention K/V stream (0 = every SM), info[2] / info[3] = split-K of
 * the E-deep / F-deep LayerNorm-fed GEMMs, info[4] = persistent-grid cap of the skinny GEMMs (0 = none),
 * info[5] = hint flags (1 K/V stream evict-first, 2 16-row tail boxes, 4 weights evict-last).  Reporting
 * only (bench.py); n = entries available in info. */
int b200_engine_decode_plan_info(const b200_engine* e, int32_t* info, int32_t n);
/* one position for every row: tokens_in [B*beam] at position pos -> argmax ids [B*beam]
 * (first index on ties, torch.argmax, model.py:233). */
int b200_engine_decode_step(b200_engine* e, const int64_t* tokens_in, int32_t pos, int64_t* next_ids,
                            void* stream);
/* the whole greedy loop without host round trips: out_tokens [B,max_len] int64 holds START, the
 * generated ids up to and including END, then PAD; out_len [B] counts START..END.  With
 * stop_check_interval = n > 0 the host checks every n steps whether every row has emitted END
 * (the reference's early exit, model.py:239-240); 0 = always run max_len-1 steps. */
int b200_engine_generate_greedy(b200_engine* e, int64_t start_id, int64_t end_id, int32_t max_len,
                                int32_t stop_check_interval, int64_t* out_tokens, int32_t* out_len,
                                void* stream);
/* beam search (the reference's is a stub, model.py:245-252; semantics in DESIGN.md): sum of
 * log-probabilities, no length penalty, finished hypotheses frozen, lower index wins ties. */
int b200_engine_generate_beam(b200_engine* e, int64_t start_id, int64_t end_id, int32_t max_len,
                              int64_t* out_tokens, int32_t* out_len, float* out_score, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_DECODER_H */
