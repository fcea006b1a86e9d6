// Decoder engine: the launch sequence of decoder.TransformerDecoder.forward (decoder.py:134-193,
// post-LN nn.TransformerDecoderLayer, torch/nn/modules/transformer.py:1143-1199), of the fused
// loss (train.py:90,327) and of their backward, expressed as calls into the sm_100a kernels.
// The engine owns no memory: parameters/gradients are caller-bound flat arenas, activations live
// in a caller-provided workspace carved by a deterministic bump plan.
#include "../../include/b200_decoder.h"
#include "attention.cuh"
#include "common.cuh"
#include "decode.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include <math.h>
#include <stdlib.h>
#include <nvtx3/nvToolsExt.h>
#include <initializer_list>
#include <string>
#include <vector>

using namespace b200;

namespace {

// ---- measured defaults (B200, BASELINE cfg2 / cfg4 shapes; profiles/r01_decode_sweeps.txt, DESIGN.md sections 3.3 and 5);
// ---- every one has an environment override, listed in INTEGRATION.md
constexpr int kBwdStageDrop = 1;         // TMA stages the backward GEMMs give up so that the bias-gradient sums fit next to them (B200_BWD_STAGE_DROP)
constexpr bool kDecSingleCta = true;     // skinny generation GEMMs as unpaired CTAs: no cluster start-up (B200_DEC_SINGLE_CTA)
constexpr int kDecFatGridPct = 70;       // % of the SMs given to the cross-attention stream when partitions run concurrently: 104 of 148 (B200_DEC_ATTN_GRID)
constexpr int kDecKvFlags = 11;          // attn_decode flags, decode.cuh: K/V evict-first + 16-row tail boxes + 3-deep rings (B200_DEC_KV_FLAGS)
// measured negative, kept switchable for the next round's re-measurement:
constexpr double kDecPrefetchMB = 0.0;   // MB (K + V) of the next layer's cross K/V warmed into L2 by the attention producer (B200_DEC_PREFETCH_MB)
constexpr bool kDecAttnDyn = false;      // atomic work counter for the attention items (B200_DEC_ATTN_DYN)
constexpr bool kDecAttnStream = false;   // cross attention of all partitions on one dedicated stream (B200_DEC_ATTN_STREAM)

// NVTX ranges (the reference has no tracing at all) around the pieces of a step / a generation call,
// switched on with B200_NVTX=1; the Python layer (engine.py, dp.py) adds the ranges of the calls it makes.
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) {
    static const bool enabled = getenv("B200_NVTX") != nullptr && atoi(getenv("B200_NVTX")) != 0;
    on = enabled;
    if (on) nvtxRangePushA(name);
  }
  ~NvtxRange() { if (on) nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct ParamInfo {
  std::string name;
  int64_t offset, numel;
};

struct LayerOff {
  int64_t sa_w, sa_b, sa_ow, sa_ob, ca_w, ca_b, ca_ow, ca_ob, l1_w, l1_b, l2_w, l2_b;
  int64_t n1_w, n1_b, n2_w, n2_b, n3_w, n3_b;
  int64_t begin;
};

struct LayerAct {   // activations of one layer kept for backward
  bf16 *qkv, *attn_o, *y1, *x1, *qc, *kvc, *attn_c, *y2, *x2, *h, *y3;
  float *lse_s, *lse_c, *mean1, *rstd1, *mean2, *rstd2, *mean3, *rstd3;
};

struct Plan {
  int B = 0, T = 0, S = 0, mem_dim = 0, training = 0;
  // rows of the decoder-side activations: B * T, or -- packed (var-len) batches -- the number of kept (non-PAD) positions,
  // sample b owning rows [cu[b], cu[b+1]) (tokenizer.py:293-313 pads every caption; packing drops those rows)
  int M = 0;
  const int32_t* cu = nullptr;     // device [B+1], caller-owned; null = regular [B, T] layout
  int64_t *ptok = nullptr, *ptgt = nullptr;   // packed token / target ids [M]
  int32_t* ppos = nullptr;                    // position of every packed row [M]
  int64_t bytes = 0;
  bf16* mem16 = nullptr;      // [B*S, mem_dim]
  bf16* memp = nullptr;       // [B*S, E] projected memory (== mem16 when no projection)
  std::vector<LayerAct> act;  // L entries when training, 1 otherwise
  std::vector<bf16*> xs;      // xs[l] = input of layer l, xs[L] = decoder output
  bf16* x_final = nullptr;    // [M, E]
  float *row_lse = nullptr, *row_loss = nullptr, *ce_scratch = nullptr, *scalars = nullptr;
  // backward temporaries
  bf16* dxd = nullptr;        // [M, E] dropout-masked copy of a LayerNorm input gradient (dropout > 0 only)
  bf16 *dlogits = nullptr, *dxa = nullptr, *dxb = nullptr, *dxc = nullptr, *dh = nullptr,
       *dqkv = nullptr, *dkv = nullptr, *dattn = nullptr, *dqc = nullptr, *dmemp16 = nullptr;
  float* dmemp = nullptr;     // [B*S, E] fp32 accumulator of d(projected memory)
};

struct Bump {
  uint8_t* base;
  int64_t off = 0;
  explicit Bump(uint8_t* b) : base(b) {}
  template <class Tp>
  Tp* take(int64_t n) {
    off = (off + 255) & ~static_cast<int64_t>(255);
    Tp* p = base ? reinterpret_cast<Tp*>(base + off) : nullptr;
    off += n * static_cast<int64_t>(sizeof(Tp));
    return p;
  }
};

}  // namespace

struct b200_engine {
  b200_engine_config cfg;
  std::vector<ParamInfo> params;
  std::vector<LayerOff> lo;
  int64_t total = 0;
  int64_t proj_w = -1, proj_b = -1, emb = -1, fc_w = -1, fc_b = -1;
  float* pf = nullptr;
  bf16* ph = nullptr;
  float* gf = nullptr;
  const float* pe = nullptr;
  // dropout (reference config.py:69 DROPOUT = 0.1): probability + caller-owned device state [seed, counter]
  float drop_p = 0.f;
  uint32_t* drop_state = nullptr;
  bool mem_bf16 = false;         // `memory` arguments are bf16 (cached frozen-encoder features) instead of fp32
  bool plan_dropout = false;     // the last training forward applied dropout (backward must regenerate the masks)
  uint8_t* ws = nullptr;
  int64_t ws_bytes = 0;
  Plan plan;   // of the last forward
  const int64_t* last_tokens = nullptr;
  const int64_t* last_targets = nullptr;
  const uint8_t* last_mem_pad = nullptr;
  long long last_ignore = 0;
  bool have_saved = false;
  // Bias-gradient column sums of the training backward run on a side stream (no shared memory, few
  // registers: they fit next to the resident GEMM CTAs and read dY while the dgrad / wgrad GEMMs stream the same
  // tiles through L2); the branch is joined back before every gradient-bucket boundary.
  cudaStream_t bias_stream = nullptr;
  cudaEvent_t bias_fork = nullptr, bias_join = nullptr;
  // B200_WGRAD_SIDE: the weight-gradient GEMMs ride the same side stream; one completion event per wgrad site of a layer
  // (+ LM head, projection) orders the main stream's next overwrite of the dY buffer that wgrad reads
  static constexpr int W_SITES = 9;
  cudaEvent_t wdone[W_SITES] = {};
  // decode state (carved from the caller's decode workspace by decode_begin)
  struct Decode {
    int B = 0, beam = 0, R = 0, S = 0, max_len = 0;
    bf16 *mem16 = nullptr, *memp = nullptr, *kv_tmp = nullptr;
    bf16 *kc = nullptr, *vc = nullptr;           // [L][B][H][S][hd] cross K/V, once per image
    bf16 *kcache[2] = {nullptr, nullptr}, *vcache[2] = {nullptr, nullptr};   // [L][R][H][max_len][hd]
    int cur = 0;                                 // which self-cache copy is live (beam ping-pong)
    bf16 *xa = nullptr, *xb = nullptr, *qkv = nullptr, *attn = nullptr, *y = nullptr, *x1 = nullptr,
         *qc = nullptr, *x2 = nullptr, *h = nullptr;
    float *scratch = nullptr, *logits = nullptr, *scores[2] = {nullptr, nullptr}, *best = nullptr;
    int64_t *ids = nullptr, *cur_tok = nullptr, *seq[2] = {nullptr, nullptr};
    int *parent = nullptr, *out_len = nullptr, *n_finished = nullptr;
    int64_t* out_tok = nullptr;    // [B, max_len] staging of the best hypotheses (stable address for the graph)
    float* out_score = nullptr;    // [B]
    unsigned char* fin[2] = {nullptr, nullptr};
    const uint8_t* mem_pad = nullptr;
    int64_t bytes = 0;
    bool ready = false;
    // Independent image ranges ("partitions"): every decode op is row-local, so each partition runs
    // its own launch chain on its own stream.  A chain is latency-bound (~70 dependent small
    // kernels per position) while the cross-attention K/V stream is HBM-bound; concurrent chains
    // fill the SMs the skinny GEMMs leave idle and keep the HBM pipe busy.
    struct Part { int b0 = 0, nb = 0; float* parts = nullptr; int* sched = nullptr; };
    int* sched = nullptr;               // [8][2] work counters of the dynamically scheduled cross attention (zero between launches)
    bool attn_dyn = false;
    std::vector<Part> part;
    int ksplit_e = 1, ksplit_f = 1;     // split-K of the E-deep / F-deep skinny GEMMs feeding a LayerNorm
    int64_t pf_bytes = 0;               // per K / V plane: head of the next layer's cross K/V warmed into L2 (0 = off)
    int kv_flags = 0;                   // 1 = K/V stream evict-first in L2, 2 = 16-row tail boxes (attn_decode flags), 4 = weights evict-last, 8 = 3-deep rings in fat CTAs
    bool single_cta = false;            // skinny GEMMs as unpaired CTAs
    int attn_fat_grid = 0;              // > 0: cross attention as <= this many one-per-SM fat CTAs (decode.cuh)
    int gemm_cap = 0;                   // > 0: persistent-grid cap of the generation GEMMs (the SMs left by the above)
    std::vector<cudaStream_t> side;     // streams of partitions 1.. (partition 0 runs on the caller's)
    std::vector<cudaEvent_t> ev;        // [0] fork, [p] join of partition p, [P] join of the attention stream
    // Optional dedicated stream for the cross-attention launches of ALL partitions (round robin): at most one
    // K/V stream runs at a time, on its SM budget (attn_fat_grid), while the other partitions' chains run on
    // the remaining SMs.  evq[p]: partition p's query is ready; eva[p]: its attention output is ready.
    bool use_att_stream = false;
    cudaStream_t att_stream = nullptr;
    std::vector<cudaEvent_t> evq, eva;
    // CUDA-graph cache of the whole generation loop (one per (workspace, shape, ids) key)
    cudaGraphExec_t graph = nullptr;
    cudaStream_t cap_stream = nullptr;
    uint64_t graph_key = 0, seen_key = 0;
  } dec;
};

namespace {

int64_t add_param(b200_engine* e, const std::string& name, int64_t numel) {
  const int64_t off = (e->total + 63) & ~static_cast<int64_t>(63);
  e->params.push_back({name, off, numel});
  e->total = off + numel;
  return off;
}

void build_plan(const b200_engine* e, Plan* pl, uint8_t* base, int B, int T, int S, int mem_dim, int training, int rows = 0) {
  const auto& c = e->cfg;
  const int64_t E = c.embed_dim, F = c.ff_dim, V = c.vocab_size, H = c.num_heads, L = c.num_layers;
  const int64_t M = rows > 0 ? rows : static_cast<int64_t>(B) * T, Ms = static_cast<int64_t>(B) * S;
  Bump b(base);
  pl->B = B; pl->T = T; pl->S = S; pl->mem_dim = mem_dim; pl->training = training; pl->M = static_cast<int>(M);
  pl->ptok = b.take<int64_t>(M); pl->ptgt = b.take<int64_t>(M); pl->ppos = b.take<int32_t>(M);
  pl->mem16 = b.take<bf16>(Ms * mem_dim);
  pl->memp = (mem_dim != E) ? b.take<bf16>(Ms * E) : pl->mem16;
  // residual stream: training keeps every layer input (xs[l]) for backward, inference ping-pongs
  pl->xs.resize(L + 1);
  if (training) {
    for (int l = 0; l <= L; ++l) pl->xs[l] = b.take<bf16>(M * E);
  } else {
    bf16* xa = b.take<bf16>(M * E);
    bf16* xb = b.take<bf16>(M * E);
    for (int l = 0; l <= L; ++l) pl->xs[l] = (l & 1) ? xb : xa;
  }
  const int slots = training ? static_cast<int>(L) : 1;
  pl->act.resize(slots);
  for (int l = 0; l < slots; ++l) {
    LayerAct& a = pl->act[l];
    a.qkv = b.take<bf16>(M * 3 * E);
    a.attn_o = b.take<bf16>(M * E);
    a.y1 = b.take<bf16>(M * E);
    a.x1 = b.take<bf16>(M * E);
    a.qc = b.take<bf16>(M * E);
    a.kvc = b.take<bf16>(Ms * 2 * E);
    a.attn_c = b.take<bf16>(M * E);
    a.y2 = b.take<bf16>(M * E);
    a.x2 = b.take<bf16>(M * E);
    a.h = b.take<bf16>(M * F);
    a.y3 = b.take<bf16>(M * E);
    a.lse_s = b.take<float>(static_cast<int64_t>(B) * H * T);
    a.lse_c = b.take<float>(static_cast<int64_t>(B) * H * T);
    a.mean1 = b.take<float>(M); a.rstd1 = b.take<float>(M);
    a.mean2 = b.take<float>(M); a.rstd2 = b.take<float>(M);
    a.mean3 = b.take<float>(M); a.rstd3 = b.take<float>(M);
  }
  pl->x_final = pl->xs[L];
  const int64_t n_tiles = (V + 127) / 128;
  pl->row_lse = b.take<float>(M);
  pl->row_loss = b.take<float>(M);
  pl->ce_scratch = b.take<float>(3 * M * n_tiles + M);
  pl->scalars = b.take<float>(16);
  if (training) {
    pl->dlogits = b.take<bf16>(M * V);
    pl->dxa = b.take<bf16>(M * E);
    pl->dxb = b.take<bf16>(M * E);
    pl->dxc = b.take<bf16>(M * E);
    pl->dxd = b.take<bf16>(M * E);
    pl->dh = b.take<bf16>(M * F);
    pl->dqkv = b.take<bf16>(M * 3 * E);
    pl->dkv = b.take<bf16>(Ms * 2 * E);
    pl->dattn = b.take<bf16>(M * E);
    pl->dqc = b.take<bf16>(M * E);
    pl->dmemp = b.take<float>(Ms * E);
    pl->dmemp16 = b.take<bf16>(Ms * E);
  }
  pl->bytes = (b.off + 255) & ~static_cast<int64_t>(255);
}

// y = x W^T + b  with optional activation / residual; W rows [row0,row0+N) of a [*,K] matrix
// y = act(x W^T + b) for a decode position (rows = hypotheses being decoded)
int linear_dec(const bf16* x, int64_t ldx, const bf16* W, const float* bias, bf16* y, int64_t ldy, int M, int N, int K,
               int act, cudaStream_t s, bool single_cta = false) {
  GemmProblem g;
  g.single_cta = single_cta;
  g.M = M; g.N = N; g.K = K;
  g.A = x; g.lda = ldx; g.B = W; g.ldb = K;
  g.D = y; g.ldd = ldy; g.bias = bias; g.act = act; g.split_k = 1; g.block_n = 128;
  return gemm_launch(g, s);
}
const DropCfg NO_DROP = DropCfg{nullptr, 0u, 0u, 1.f};
int linear_fwd(const bf16* x, int64_t ldx, const bf16* W, const float* bias, bf16* y, int64_t ldy, int M,
               int N, int K, int act, const bf16* residual, int64_t ldr, cudaStream_t s, DropCfg drop = NO_DROP) {
  GemmProblem g;
  g.drop = drop;
  g.M = M; g.N = N; g.K = K;
  g.A = x; g.lda = ldx; g.B = W; g.ldb = K;
  g.D = y; g.ldd = ldy; g.bias = bias; g.act = act; g.residual = residual; g.ldr = ldr;
  g.split_k = 1;
  return gemm_launch(g, s);
}
// dx = dy W (+ residual) (* relu mask);  W is [N_out, K_in] as stored
int linear_dgrad(const bf16* dy, int64_t lddy, const bf16* W, int N_out, int K_in, bf16* dx, int64_t lddx,
                 int M, const bf16* residual, int64_t ldr, const bf16* relu_mask, int64_t ldm, cudaStream_t s,
                 float mask_scale = 1.f) {
  GemmProblem g;
  g.mask_scale = mask_scale;
  g.M = M; g.N = K_in; g.K = N_out;
  g.A = dy; g.lda = lddy; g.B = W; g.ldb = K_in; g.b_mn = true;
  g.D = dx; g.ldd = lddx; g.residual = residual; g.ldr = ldr; g.relu_mask = relu_mask; g.ldm = ldm;
  g.split_k = 1;
  return gemm_launch(g, s);
}
// dW[N_out,K_in] += dy^T x ; db[N_out] += colsum(dy)
// side != null: the bias sum runs on that stream behind `fork` (dy is complete at this point of s); the caller
// joins the side stream before dy is overwritten and before the gradients are consumed.  done != null: the wgrad GEMM
// rides the side stream too (dW is needed only at the bucket boundary, so it fills the SMs the main stream's
// HBM-bound kernels and GEMM tails leave idle) and `done` is recorded behind it.
int linear_wgrad(const bf16* dy, int64_t lddy, const bf16* x, int64_t ldx, float* dW, float* db, int M,
                 int N_out, int K_in, cudaStream_t s, cudaStream_t side = nullptr, cudaEvent_t fork = nullptr,
                 bool* side_used = nullptr, cudaEvent_t done = nullptr) {
  GemmProblem g;
  g.M = N_out; g.N = K_in; g.K = M;
  g.A = dy; g.lda = lddy; g.a_mn = true; g.B = x; g.ldb = ldx; g.b_mn = true;
  g.D = dW; g.ldd = K_in; g.d_fp32 = true; g.accumulate = true; g.split_k = 0;
  if (side && (db || done)) {
    B200_CHECK_CUDA(cudaEventRecord(fork, s));
    B200_CHECK_CUDA(cudaStreamWaitEvent(side, fork, 0));
    if (side_used) *side_used = true;
    if (db) { if (int rc = colsum(dy, lddy, db, M, N_out, side)) return rc; }
    if (!done) return gemm_launch(g, s);
    if (int rc = gemm_launch(g, side)) return rc;
    B200_CHECK_CUDA(cudaEventRecord(done, side));
    return 0;
  }
  if (int rc = gemm_launch(g, s)) return rc;
  if (db) return colsum(dy, lddy, db, M, N_out, s);
  return 0;
}

#define RC(expr) do { if (int _rc = (expr)) return _rc; } while (0)

DropCfg drop_site(const b200_engine* e, int id) {
  DropCfg d;
  d.state = e->drop_state;
  d.site = static_cast<uint32_t>(id);
  d.thr = static_cast<uint32_t>(e->drop_p * 65536.0f + 0.5f);
  d.scale = 1.0f / (1.0f - e->drop_p);
  return d;
}

int run_forward(b200_engine* e, const int64_t* tokens, const float* memory, const uint8_t* mem_pad, int B,
                int T, int S, int mem_dim, int training, cudaStream_t s, const int32_t* cu = nullptr, int rows = 0,
                const int64_t* targets = nullptr) {
  const auto& c = e->cfg;
  const int E = c.embed_dim, F = c.ff_dim, H = c.num_heads, L = c.num_layers, hd = E / H;
  B200_REQUIRE(e->pf && e->ph, "engine: parameters not bound");
  B200_REQUIRE(B > 0 && T > 0 && S > 0, "engine: empty batch (B=%d T=%d S=%d)", B, T, S);
  B200_REQUIRE(T <= c.max_seq_len, "engine: T=%d exceeds max_seq_len=%d (positional table too short, decoder.py:71)", T, c.max_seq_len);
  B200_REQUIRE(mem_dim == E || (mem_dim == c.enc_dim && e->proj_w >= 0), "engine: memory width %d matches neither embed_dim %d nor enc_dim %d", mem_dim, E, c.enc_dim);
  B200_REQUIRE(cu == nullptr || (rows > 0 && rows <= B * T), "engine: packed batch with %d rows (B*T = %d)", rows, B * T);
  Plan probe;
  build_plan(e, &probe, nullptr, B, T, S, mem_dim, training, cu ? rows : 0);
  B200_REQUIRE(e->ws && probe.bytes <= e->ws_bytes, "engine: workspace too small (%lld needed, %lld bound)", (long long)probe.bytes, (long long)e->ws_bytes);
  build_plan(e, &e->plan, e->ws, B, T, S, mem_dim, training, cu ? rows : 0);
  Plan& pl = e->plan;
  pl.cu = cu;
  const int M = pl.M, Ms = B * S;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  NvtxRange nvtx_fwd("b200.decoder.forward");

  // dropout only in training mode (model.train(), train.py:64); eval / generation never drop
  const bool dropping = training && e->drop_p > 0.f;
  B200_REQUIRE(!dropping || e->drop_state, "engine: dropout enabled without a device state buffer");
  e->plan_dropout = dropping;
  if (dropping) RC(drop_advance(e->drop_state, s));
  auto site = [&](int id) { return dropping ? drop_site(e, id) : NO_DROP; };

  if (e->mem_bf16) {
    // bf16 features are consumed in place (no cast pass); the caller keeps them alive and unchanged until backward is done
    B200_REQUIRE((reinterpret_cast<uintptr_t>(memory) & 15) == 0, "engine: bf16 memory must be 16-byte aligned");
    const bool same = pl.memp == pl.mem16;
    pl.mem16 = reinterpret_cast<bf16*>(const_cast<float*>(memory));
    if (same) pl.memp = pl.mem16;
  } else {
    RC(cast_f32_to_bf16(memory, pl.mem16, static_cast<long long>(Ms) * mem_dim, s));
  }
  if (mem_dim != E)
    RC(linear_fwd(pl.mem16, mem_dim, e->ph + e->proj_w, e->pf + e->proj_b, pl.memp, E, Ms, E, mem_dim, 0, nullptr, 0, s));

  if (cu) {
    // gather the kept positions: packed token / target ids and the position of every packed row
    RC(pack_rows(tokens, targets, cu, B, T, pl.ptok, pl.ptgt, pl.ppos, s));
    RC(embed_pe_fwd(pl.ptok, e->pf + e->emb, e->pe, pl.xs[0], M, 1, E, c.vocab_size, sqrtf(static_cast<float>(E)), s, 0, site(0), pl.ppos));
  } else {
    RC(embed_pe_fwd(tokens, e->pf + e->emb, e->pe, pl.xs[0], B, T, E, c.vocab_size, sqrtf(static_cast<float>(E)), s, 0, site(0)));
  }

  for (int l = 0; l < L; ++l) {
    NvtxRange nvtx_layer("b200.decoder.forward.layer");
    const LayerOff& o = e->lo[l];
    LayerAct& a = pl.act[training ? l : 0];
    const bf16* x = pl.xs[l];
    bf16* x_out = pl.xs[l + 1];
    // --- self-attention block
    RC(linear_fwd(x, E, e->ph + o.sa_w, e->pf + o.sa_b, a.qkv, 3 * E, M, 3 * E, E, 0, nullptr, 0, s));
    AttnArgs sa;
    sa.q = a.qkv; sa.k = a.qkv + E; sa.v = a.qkv + 2 * E;
    sa.q_bs = sa.k_bs = sa.v_bs = static_cast<long long>(T) * 3 * E; sa.q_ts = sa.k_ts = sa.v_ts = 3 * E;
    sa.o = a.attn_o; sa.o_bs = static_cast<long long>(T) * E; sa.o_ts = E;
    sa.lse = a.lse_s; sa.B = B; sa.H = H; sa.Tq = T; sa.Tk = T; sa.hd = hd; sa.causal = 1;
    sa.key_tokens = cu ? nullptr : tokens; sa.pad_idx = c.pad_idx; sa.scale = scale; sa.drop = site(1 + 6 * l + 0);
    sa.cu_q = sa.cu_k = cu; sa.total_q = sa.total_k = M;      // packed: no PAD keys exist, every sample has its own length
    RC(attn_fwd(sa, s));
    RC(linear_fwd(a.attn_o, E, e->ph + o.sa_ow, e->pf + o.sa_ob, a.y1, E, M, E, E, 0, x, E, s, site(1 + 6 * l + 1)));
    RC(layernorm_fwd(a.y1, e->pf + o.n1_w, e->pf + o.n1_b, a.x1, a.mean1, a.rstd1, M, E, c.ln_eps, s));
    // --- cross-attention block
    RC(linear_fwd(a.x1, E, e->ph + o.ca_w, e->pf + o.ca_b, a.qc, E, M, E, E, 0, nullptr, 0, s));
    RC(linear_fwd(pl.memp, E, e->ph + o.ca_w + static_cast<int64_t>(E) * E, e->pf + o.ca_b + E, a.kvc, 2 * E, Ms, 2 * E, E, 0, nullptr, 0, s));
    AttnArgs ca;
    ca.q = a.qc; ca.q_bs = static_cast<long long>(T) * E; ca.q_ts = E;
    ca.k = a.kvc; ca.v = a.kvc + E; ca.k_bs = ca.v_bs = static_cast<long long>(S) * 2 * E; ca.k_ts = ca.v_ts = 2 * E;
    ca.o = a.attn_c; ca.o_bs = static_cast<long long>(T) * E; ca.o_ts = E;
    ca.lse = a.lse_c; ca.B = B; ca.H = H; ca.Tq = T; ca.Tk = S; ca.hd = hd; ca.causal = 0;
    ca.key_pad_mask = mem_pad; ca.scale = scale; ca.drop = site(1 + 6 * l + 2);
    ca.cu_q = cu; ca.total_q = M;
    RC(attn_fwd(ca, s));
    RC(linear_fwd(a.attn_c, E, e->ph + o.ca_ow, e->pf + o.ca_ob, a.y2, E, M, E, E, 0, a.x1, E, s, site(1 + 6 * l + 3)));
    RC(layernorm_fwd(a.y2, e->pf + o.n2_w, e->pf + o.n2_b, a.x2, a.mean2, a.rstd2, M, E, c.ln_eps, s));
    // --- feed-forward block
    RC(linear_fwd(a.x2, E, e->ph + o.l1_w, e->pf + o.l1_b, a.h, F, M, F, E, c.act, nullptr, 0, s, site(1 + 6 * l + 4)));
    RC(linear_fwd(a.h, F, e->ph + o.l2_w, e->pf + o.l2_b, a.y3, E, M, E, F, 0, a.x2, E, s, site(1 + 6 * l + 5)));
    RC(layernorm_fwd(a.y3, e->pf + o.n3_w, e->pf + o.n3_b, x_out, a.mean3, a.rstd3, M, E, c.ln_eps, s));
  }
  return 0;
}

// Backward in L+2 parts, in the order gradients become final: part 0 = LM head (fc_out), part k in
// [1, L] = decoder layer L-k, part L+1 = embedding (+ projection).  Parts [first, last] are run;
// the data-parallel driver runs one part per gradient bucket and all-reduces in between.
int run_backward(b200_engine* e, bool have_dlogits, float* dmemory, void* const* events, int n_events, cudaStream_t s,
                 int first_part = 0, int last_part = 1 << 30) {
  const auto& c = e->cfg;
  const int E = c.embed_dim, F = c.ff_dim, H = c.num_heads, L = c.num_layers, hd = E / H, V = c.vocab_size;
  Plan& pl = e->plan;
  B200_REQUIRE(e->have_saved && pl.training, "engine: backward needs a preceding training forward");
  B200_REQUIRE(e->gf, "engine: gradient arena not bound");
  B200_REQUIRE(c.act == B200_ACT_RELU, "engine: backward is implemented for the reference activation (ReLU) only");
  const int B = pl.B, T = pl.T, S = pl.S;
  const int M = pl.M, Ms = B * S;
  const int32_t* cu = pl.cu;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  const bool need_dmemp = (pl.mem_dim != E) || dmemory != nullptr;
  B200_REQUIRE(!(dmemory && pl.mem_dim != E), "engine: dmemory is only available when memory is already embed_dim wide");
  (void)have_dlogits;
  const bool dropping = e->plan_dropout;
  auto site = [&](int id) { return dropping ? drop_site(e, id) : NO_DROP; };
  const float keep_scale = dropping ? 1.0f / (1.0f - e->drop_p) : 1.0f;
  const bool bias_inline = getenv("B200_BIAS_INLINE") != nullptr;      // read per call: A/B switch and bench.py's roofline pass
  cudaStream_t side = nullptr;
  if (!bias_inline) {
    if (!e->bias_stream) B200_CHECK_CUDA(cudaStreamCreateWithFlags(&e->bias_stream, cudaStreamNonBlocking));
    if (!e->bias_fork) B200_CHECK_CUDA(cudaEventCreateWithFlags(&e->bias_fork, cudaEventDisableTiming));
    if (!e->bias_join) B200_CHECK_CUDA(cudaEventCreateWithFlags(&e->bias_join, cudaEventDisableTiming));
    for (auto& w : e->wdone) if (!w) B200_CHECK_CUDA(cudaEventCreateWithFlags(&w, cudaEventDisableTiming));
    side = e->bias_stream;
  }
  // one pipeline stage less while bias sums may be co-resident (B200_BWD_STAGE_DROP overrides)
  static const int bwd_drop = getenv("B200_BWD_STAGE_DROP") ? atoi(getenv("B200_BWD_STAGE_DROP")) : kBwdStageDrop;
  GemmStageCap stage_cap(side ? bwd_drop : 0);
  bool side_used = false;
  // B200_WGRAD_SIDE=1 (read per call: A/B switch): the wgrad GEMMs ride the side stream as well; wpend[site] = a wgrad of
  // that site is in flight, wwait(site) orders the main stream behind it (and, the side stream being in order, behind
  // everything launched on it before)
  enum { W_L2 = 0, W_L1, W_CAO, W_CAQ, W_CAKV, W_SAO, W_SAQKV, W_FC, W_PROJ };
  const char* wside_env = getenv("B200_WGRAD_SIDE");
  const bool wside = side != nullptr && wside_env != nullptr && atoi(wside_env) != 0;
  bool wpend[b200_engine::W_SITES] = {};
  // every return path below goes through join_side (a captured graph must not be left forked)
  auto join_side = [&](void) {
    if (side && side_used) {
      cudaEventRecord(e->bias_join, side);
      cudaStreamWaitEvent(s, e->bias_join, 0);
      side_used = false;
      for (auto& w : wpend) w = false;
    }
  };
  auto wwait = [&](int site_id) {
    if (wpend[site_id]) {
      cudaStreamWaitEvent(s, e->wdone[site_id], 0);
      wpend[site_id] = false;
    }
  };
  auto wgrad = [&](int site_id, const bf16* dy, int64_t lddy, const bf16* x, int64_t ldx, float* dW, float* db, int Mr,
                   int N_out, int K_in) -> int {
    if (!wside) return linear_wgrad(dy, lddy, x, ldx, dW, db, Mr, N_out, K_in, s, side, e->bias_fork, &side_used);
    wpend[site_id] = true;
    return linear_wgrad(dy, lddy, x, ldx, dW, db, Mr, N_out, K_in, s, side, e->bias_fork, &side_used, e->wdone[site_id]);
  };
  struct Joiner { decltype(join_side)& j; ~Joiner() { j(); } } joiner{join_side};
  // Joins: (1) before a gradient-bucket event is recorded (data-parallel runs consume the bucket's bias gradients
  // behind it), (2) in every layer right before the FFN dgrad rewrites pl.dh, the first buffer a pending sum of
  // the previous layer may still be reading (by then those sums are long done: the wait is free), (3) at the end.
  int ev = 0;
  auto mark = [&](void) {
    if (events && ev < n_events && events[ev]) {
      join_side();
      cudaEventRecord(static_cast<cudaEvent_t>(events[ev]), s);
    }
    ++ev;
  };

  NvtxRange nvtx_bwd("b200.decoder.backward");
  bf16* dx = pl.dxa;      // gradient of the residual stream at every part boundary lives in dxa
  if (first_part <= 0) {
    NvtxRange nvtx_part("b200.decoder.backward.lm_head");
    if (need_dmemp) B200_CHECK_CUDA(cudaMemsetAsync(pl.dmemp, 0, static_cast<size_t>(Ms) * E * sizeof(float), s));
    // --- LM head
    RC(wgrad(W_FC, pl.dlogits, V, pl.x_final, E, e->gf + e->fc_w, e->gf + e->fc_b, M, V, E));
    RC(linear_dgrad(pl.dlogits, V, e->ph + e->fc_w, V, E, dx, E, M, nullptr, 0, nullptr, 0, s));
    mark();
  }

  bf16* spare1 = pl.dxb;
  bf16* spare2 = pl.dxc;
  for (int l = L - 1; l >= 0; --l) {
    const int part = L - l;
    if (part < first_part || part > last_part) continue;
    NvtxRange nvtx_part("b200.decoder.backward.layer");
    const LayerOff& o = e->lo[l];
    LayerAct& a = pl.act[l];
    float* g = e->gf;
    // LN3
    // With dropout, y = x + dropout(z): the LayerNorm backward emits the gradient of y (residual
    // branch) and, masked and rescaled, the gradient of z (operand of the Linear's dgrad / wgrad and
    // of its bias sum).  Without dropout the two coincide and dz* alias dy*.
    bf16* dy3 = spare1;
    bf16* dz3 = dropping ? pl.dxd : dy3;
    // (wside waits: the buffer a kernel is about to overwrite may still be read by a wgrad on the side stream -- the
    // previous layer's for the per-site buffers, this layer's for the rotating dx buffers; pl.dxd carries all three dz
    // of a layer when dropout is on)
    if (dropping) wwait(W_SAO);
    RC(layernorm_bwd(dx, a.y3, e->pf + o.n3_w, a.mean3, a.rstd3, dy3, g + o.n3_w, g + o.n3_b, g + o.l2_b, M, E, s,
                     dropping ? dz3 : nullptr, site(1 + 6 * l + 5)));
    // FFN (linear2's bias gradient = column sums of dz3, produced by the LayerNorm backward above)
    RC(wgrad(W_L2, dz3, E, a.h, F, g + o.l2_w, nullptr, M, E, F));
    // h = dropout(relu(.)) is positive exactly where the unit is active AND kept
    if (wside) wwait(W_L1); else join_side();
    RC(linear_dgrad(dz3, E, e->ph + o.l2_w, E, F, pl.dh, F, M, nullptr, 0, a.h, F, s, keep_scale));
    RC(wgrad(W_L1, pl.dh, F, a.x2, E, g + o.l1_w, g + o.l1_b, M, F, E));
    bf16* dx2 = spare2;
    wwait(W_SAO);            // dz1 of the previous layer lives in spare2
    RC(linear_dgrad(pl.dh, F, e->ph + o.l1_w, F, E, dx2, E, M, dy3, E, nullptr, 0, s));
    // LN2
    bf16* dy2 = dx;   // dx (grad of layer output) is dead now
    bf16* dz2 = dropping ? pl.dxd : dy2;      // dz3 is dead (its last reader, the FFN dgrad above, is ordered before)
    if (dropping) wwait(W_L2);
    RC(layernorm_bwd(dx2, a.y2, e->pf + o.n2_w, a.mean2, a.rstd2, dy2, g + o.n2_w, g + o.n2_b, g + o.ca_ob, M, E, s,
                     dropping ? dz2 : nullptr, site(1 + 6 * l + 3)));
    // cross-attention
    RC(wgrad(W_CAO, dz2, E, a.attn_c, E, g + o.ca_ow, nullptr, M, E, E));
    RC(linear_dgrad(dz2, E, e->ph + o.ca_ow, E, E, pl.dattn, E, M, nullptr, 0, nullptr, 0, s));
    AttnArgs ca;
    ca.q = a.qc; ca.q_bs = static_cast<long long>(T) * E; ca.q_ts = E;
    ca.k = a.kvc; ca.v = a.kvc + E; ca.k_bs = ca.v_bs = static_cast<long long>(S) * 2 * E; ca.k_ts = ca.v_ts = 2 * E;
    ca.o = a.attn_c; ca.o_bs = static_cast<long long>(T) * E; ca.o_ts = E;
    ca.lse = a.lse_c; ca.B = B; ca.H = H; ca.Tq = T; ca.Tk = S; ca.hd = hd; ca.causal = 0;
    ca.key_pad_mask = e->last_mem_pad; ca.scale = scale; ca.drop = site(1 + 6 * l + 2);
    ca.cu_q = cu; ca.total_q = M;
    AttnGrads cg;
    cg.d_o = pl.dattn; cg.do_bs = static_cast<long long>(T) * E; cg.do_ts = E;
    cg.dq = pl.dqc; cg.dq_bs = static_cast<long long>(T) * E; cg.dq_ts = E;
    cg.dk = pl.dkv; cg.dv = pl.dkv + E; cg.dk_bs = cg.dv_bs = static_cast<long long>(S) * 2 * E; cg.dk_ts = cg.dv_ts = 2 * E;
    wwait(W_CAQ);
    wwait(W_CAKV);           // the previous layer's readers of pl.dqc / pl.dkv
    RC(attn_bwd(ca, cg, s));
    RC(wgrad(W_CAQ, pl.dqc, E, a.x1, E, g + o.ca_w, g + o.ca_b, M, E, E));
    RC(wgrad(W_CAKV, pl.dkv, 2 * E, pl.memp, E, g + o.ca_w + static_cast<int64_t>(E) * E, g + o.ca_b + E, Ms, 2 * E, E));
    if (need_dmemp) {
      GemmProblem gp;
      gp.M = Ms; gp.N = E; gp.K = 2 * E;
      gp.A = pl.dkv; gp.lda = 2 * E; gp.B = e->ph + o.ca_w + static_cast<int64_t>(E) * E; gp.ldb = E; gp.b_mn = true;
      gp.D = pl.dmemp; gp.ldd = E; gp.d_fp32 = true; gp.accumulate = true; gp.split_k = 1;
      RC(gemm_launch(gp, s));
    }
    bf16* dx1 = spare1;   // dy3 is dead
    wwait(W_L2);             // ... once this layer's linear2 wgrad has read it
    RC(linear_dgrad(pl.dqc, E, e->ph + o.ca_w, E, E, dx1, E, M, dy2, E, nullptr, 0, s));
    // LN1
    bf16* dy1 = spare2;   // dx2 is dead
    bf16* dz1 = dropping ? pl.dxd : dy1;      // dz2 is dead
    if (dropping) wwait(W_CAO);
    RC(layernorm_bwd(dx1, a.y1, e->pf + o.n1_w, a.mean1, a.rstd1, dy1, g + o.n1_w, g + o.n1_b, g + o.sa_ob, M, E, s,
                     dropping ? dz1 : nullptr, site(1 + 6 * l + 1)));
    // self-attention
    RC(wgrad(W_SAO, dz1, E, a.attn_o, E, g + o.sa_ow, nullptr, M, E, E));
    RC(linear_dgrad(dz1, E, e->ph + o.sa_ow, E, E, pl.dattn, E, M, nullptr, 0, nullptr, 0, s));
    AttnArgs sa;
    sa.q = a.qkv; sa.k = a.qkv + E; sa.v = a.qkv + 2 * E;
    sa.q_bs = sa.k_bs = sa.v_bs = static_cast<long long>(T) * 3 * E; sa.q_ts = sa.k_ts = sa.v_ts = 3 * E;
    sa.o = a.attn_o; sa.o_bs = static_cast<long long>(T) * E; sa.o_ts = E;
    sa.lse = a.lse_s; sa.B = B; sa.H = H; sa.Tq = T; sa.Tk = T; sa.hd = hd; sa.causal = 1;
    sa.key_tokens = cu ? nullptr : e->last_tokens; sa.pad_idx = c.pad_idx; sa.scale = scale; sa.drop = site(1 + 6 * l + 0);
    sa.cu_q = sa.cu_k = cu; sa.total_q = sa.total_k = M;
    AttnGrads sg;
    sg.d_o = pl.dattn; sg.do_bs = static_cast<long long>(T) * E; sg.do_ts = E;
    sg.dq = pl.dqkv; sg.dk = pl.dqkv + E; sg.dv = pl.dqkv + 2 * E;
    sg.dq_bs = sg.dk_bs = sg.dv_bs = static_cast<long long>(T) * 3 * E; sg.dq_ts = sg.dk_ts = sg.dv_ts = 3 * E;
    wwait(W_SAQKV);          // the previous layer's reader of pl.dqkv
    RC(attn_bwd(sa, sg, s));
    RC(wgrad(W_SAQKV, pl.dqkv, 3 * E, pl.xs[l], E, g + o.sa_w, g + o.sa_b, M, 3 * E, E));
    bf16* dx_in = dx;     // dy2 is dead
    wwait(W_CAO);            // ... once this layer's cross out-projection wgrad has read it
    RC(linear_dgrad(pl.dqkv, 3 * E, e->ph + o.sa_w, 3 * E, E, dx_in, E, M, dy1, E, nullptr, 0, s));
    dx = dx_in;
    mark();
  }
  if (last_part < L + 1) return 0;
  if (cu) RC(embed_bwd(pl.ptok, dx, e->gf + e->emb, M, 1, E, V, c.pad_idx, sqrtf(static_cast<float>(E)), s, site(0)));
  else RC(embed_bwd(e->last_tokens, dx, e->gf + e->emb, B, T, E, V, c.pad_idx, sqrtf(static_cast<float>(E)), s, site(0)));
  if (pl.mem_dim != E) {
    RC(cast_f32_to_bf16(pl.dmemp, pl.dmemp16, static_cast<long long>(Ms) * E, s));
    RC(wgrad(W_PROJ, pl.dmemp16, E, pl.mem16, pl.mem_dim, e->gf + e->proj_w, e->gf + e->proj_b, Ms, E, pl.mem_dim));
  } else if (dmemory) {
    B200_CHECK_CUDA(cudaMemcpyAsync(dmemory, pl.dmemp, static_cast<size_t>(Ms) * E * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  mark();
  return 0;
}


void build_decode_plan(const b200_engine* e, b200_engine::Decode* d, uint8_t* base, int B, int beam, int S,
                       int mem_dim, int max_len) {
  const auto& c = e->cfg;
  const int64_t E = c.embed_dim, F = c.ff_dim, V = c.vocab_size, L = c.num_layers;
  const int64_t R = static_cast<int64_t>(B) * beam, Ms = static_cast<int64_t>(B) * S;
  Bump b(base);
  d->B = B; d->beam = beam; d->R = static_cast<int>(R); d->S = S; d->max_len = max_len;
  d->mem16 = b.take<bf16>(Ms * mem_dim);
  d->memp = (mem_dim != E) ? b.take<bf16>(Ms * E) : d->mem16;
  d->kv_tmp = b.take<bf16>(Ms * 2 * E);
  d->kc = b.take<bf16>(L * Ms * E);
  d->vc = b.take<bf16>(L * Ms * E);
  const int copies = beam > 1 ? 2 : 1;
  for (int i = 0; i < 2; ++i) {
    d->kcache[i] = i < copies ? b.take<bf16>(L * R * max_len * E) : nullptr;
    d->vcache[i] = i < copies ? b.take<bf16>(L * R * max_len * E) : nullptr;
    d->seq[i] = b.take<int64_t>(R * max_len);
    d->scores[i] = b.take<float>(R);
    d->fin[i] = b.take<unsigned char>(R);
  }
  d->xa = b.take<bf16>(R * E); d->xb = b.take<bf16>(R * E);
  d->qkv = b.take<bf16>(R * 3 * E); d->attn = b.take<bf16>(R * E); d->y = b.take<bf16>(R * E);
  d->x1 = b.take<bf16>(R * E); d->qc = b.take<bf16>(R * E); d->x2 = b.take<bf16>(R * E);
  d->h = b.take<bf16>(R * F);
  d->scratch = b.take<float>(2 * R * ((V + 127) / 128));
  d->logits = beam > 1 ? b.take<float>(R * V) : nullptr;
  d->best = b.take<float>(R);
  d->ids = b.take<int64_t>(R); d->cur_tok = b.take<int64_t>(R);
  d->parent = b.take<int>(R); d->out_len = b.take<int>(R); d->n_finished = b.take<int>(4);
  d->out_tok = b.take<int64_t>(static_cast<int64_t>(B) * max_len); d->out_score = b.take<float>(B);
  // partitions: whole 128-row GEMM tiles where possible
  const char* forced_env = getenv("B200_DECODE_PARTS");   // read per plan: tests compare partitionings
  const int forced = forced_env ? atoi(forced_env) : 0;
  // Measured on B200 (cfg2, 512 images, greedy; profiles/r01_decode_sweeps.txt): a position is a chain of ~75
  // dependent launches (latency-bound, few SMs) around six HBM-bound cross-attention streams.  Four concurrent
  // partitions of 128 rows overlap one partition's stream -- run as fat CTAs on 104 SMs -- with the other
  // partitions' chains on the remaining SMs: 32.4 vs 39.1 ms per 512 captions.  Plain small-CTA attention grids
  // fill every SM and serialise the partitions again (+3 % only).  Beam search (4x the rows per image, cache
  // re-indexing) is not chain-bound and loses 5-7 % when partitioned, small K/V streams have nothing to overlap.
  const bool stream_heavy = static_cast<double>(B) * S * E * 4.0 >= 128.0 * 1024 * 1024;   // K+V bytes of one layer
  int P = forced > 0 ? forced : ((beam == 1 && B >= 256 && stream_heavy) ? 4 : 1);
  if (P > B) P = B;
  if (P > 8) P = 8;
  int per = (B + P - 1) / P;
  const int gran = beam <= 128 ? (128 / beam > 0 ? 128 / beam : 1) : 1;   // images per 128 rows
  if (per > gran) per = (per + gran - 1) / gran * gran;
  d->part.clear();
  for (int b0 = 0; b0 < B; b0 += per) {
    b200_engine::Decode::Part pt;
    pt.b0 = b0;
    pt.nb = (B - b0 < per) ? (B - b0) : per;
    d->part.push_back(pt);
  }
  // L2 warm-up of the next layer's cross K/V stream (decode.cu: DecAttnArgs::pf_*): total MB over K and V,
  // shared by the partitions.  Read per plan so that a sweep can compare settings in one process.
  {
    const char* pf_env = getenv("B200_DEC_PREFETCH_MB");
    const double mb = pf_env ? atof(pf_env) : kDecPrefetchMB;
    const int64_t plane = static_cast<int64_t>(per) * S * E * 2;          // bytes of one partition's K (or V) plane of a layer
    int64_t want = static_cast<int64_t>(mb * 1024.0 * 1024.0 / 2.0 / static_cast<double>(d->part.size() ? d->part.size() : 1));
    if (want > plane) want = plane;
    d->pf_bytes = want > 0 ? (want & ~static_cast<int64_t>(4095)) : 0;
    const char* fl_env = getenv("B200_DEC_KV_FLAGS");
    d->kv_flags = fl_env ? (atoi(fl_env) & 15) : kDecKvFlags;
  }
  // split-K of the LayerNorm-fed GEMMs: enough CTAs to cover the SMs
  const int rows_p = per * beam;
  const int tiles = ((rows_p + 127) / 128) * static_cast<int>((E + 127) / 128);
  auto pick = [&](int64_t K) {
    const int kb = static_cast<int>((K + 63) / 64);
    int sk = 148 / (tiles > 0 ? tiles : 1);
    // >= 6 k-blocks per split and at most 4 slabs: beyond that the LayerNorm-side reduction costs more than the
    // K loop saves (sweep on B200: 2 / 4 splits for K = 768 / 3072 beat 3 / 6 by 3 %, 3 / 8 by 4 %)
    if (sk > kb / 6) sk = kb / 6;
    if (sk > 4) sk = 4;
    if (sk < 1) sk = 1;
    return gemm_effective_splits(static_cast<int>(K), sk);
  };
  d->ksplit_e = pick(E);
  d->ksplit_f = pick(F);
  if (const char* v = getenv("B200_DEC_KSPLIT_E")) { if (atoi(v) > 0) d->ksplit_e = gemm_effective_splits(static_cast<int>(E), atoi(v)); }
  if (const char* v = getenv("B200_DEC_KSPLIT_F")) { if (atoi(v) > 0) d->ksplit_f = gemm_effective_splits(static_cast<int>(F), atoi(v)); }
  {
    const char* v = getenv("B200_DEC_SINGLE_CTA");
    d->single_cta = v ? (atoi(v) != 0) : kDecSingleCta;
    const char* g = getenv("B200_DEC_ATTN_GRID");
    d->attn_fat_grid = g ? atoi(g) : (d->part.size() > 1 ? (device_sm_count() * kDecFatGridPct + 50) / 100 : 0);
    if (d->attn_fat_grid < 0) d->attn_fat_grid = 0;
    if (d->attn_fat_grid >= device_sm_count()) d->attn_fat_grid = device_sm_count();
    const char* as = getenv("B200_DEC_ATTN_STREAM");
    d->use_att_stream = d->part.size() > 1 && (as ? atoi(as) != 0 : kDecAttnStream);
    const char* gc = getenv("B200_DEC_GEMM_CTAS");
    d->gemm_cap = gc ? atoi(gc) : (d->attn_fat_grid > 0 ? device_sm_count() - d->attn_fat_grid : 0);
    if (d->gemm_cap < 0) d->gemm_cap = 0;
  }
  const int ks_max = d->ksplit_e > d->ksplit_f ? d->ksplit_e : d->ksplit_f;
  for (auto& pt : d->part) {
    const int64_t mpad = (static_cast<int64_t>(pt.nb) * beam + 127) / 128 * 128;
    pt.parts = b.take<float>(ks_max * mpad * E);
  }
  d->sched = b.take<int>(16);
  {
    const char* dy = getenv("B200_DEC_ATTN_DYN");
    d->attn_dyn = dy ? (atoi(dy) != 0) : kDecAttnDyn;
    for (size_t p = 0; p < d->part.size(); ++p) d->part[p].sched = (d->attn_dyn && d->sched) ? d->sched + 2 * p : nullptr;
  }
  d->bytes = (b.off + 255) & ~static_cast<int64_t>(255);
}

// y = LayerNorm(x W^T + b + residual) for the skinny decode GEMMs: split-K fp32 partial slabs,
// summed together with bias and residual inside the LayerNorm kernel
int linear_ln_fwd(const bf16* x, int64_t ldx, const bf16* W, const float* bias, const bf16* residual,
                  const float* gamma, const float* beta, float* parts, int split, bf16* y, int M, int N, int K,
                  float eps, cudaStream_t s, bool single_cta = false) {
  GemmProblem g;
  g.single_cta = single_cta;
  g.M = M; g.N = N; g.K = K;
  g.A = x; g.lda = ldx; g.B = W; g.ldb = K;
  g.D = parts; g.ldd = N; g.d_fp32 = true; g.partials = true; g.split_k = split; g.block_n = 128;
  RC(gemm_launch(g, s));
  const long long mpad = (static_cast<long long>(M) + 127) / 128 * 128;
  return layernorm_reduce_fwd(parts, gemm_effective_splits(K, split), mpad * N, N, bias, residual, N, gamma, beta, y, M, N, eps, s);
}

// ---- one decode position, cut into the pieces between which partitions interleave ----
struct DecPart {          // per-partition views of the decode workspace
  const b200_engine::Decode::Part* pt;
  int64_t r0; int R, B;
  bf16 *xa, *xb, *qkv, *attn, *x1, *qc, *x2, *h;
  const uint8_t* mem_pad;
  int64_t self_off, cross_off, pf_bytes;
};

DecPart dec_part(const b200_engine* e, const b200_engine::Decode::Part& pt) {
  const auto& d = e->dec;
  const int64_t E = e->cfg.embed_dim, F = e->cfg.ff_dim;
  DecPart v;
  v.pt = &pt;
  v.r0 = static_cast<int64_t>(pt.b0) * d.beam;
  v.R = pt.nb * d.beam; v.B = pt.nb;
  v.xa = d.xa + v.r0 * E; v.xb = d.xb + v.r0 * E;
  v.qkv = d.qkv + v.r0 * 3 * E; v.attn = d.attn + v.r0 * E;
  v.x1 = d.x1 + v.r0 * E; v.qc = d.qc + v.r0 * E; v.x2 = d.x2 + v.r0 * E;
  v.h = d.h + v.r0 * F;
  v.mem_pad = d.mem_pad ? d.mem_pad + static_cast<int64_t>(pt.b0) * d.S : nullptr;
  v.self_off = v.r0 * d.max_len * E;
  v.cross_off = static_cast<int64_t>(pt.b0) * d.S * E;
  const int64_t plane_bytes = static_cast<int64_t>(pt.nb) * d.S * E * 2;
  v.pf_bytes = d.pf_bytes < plane_bytes ? d.pf_bytes : (plane_bytes & ~static_cast<int64_t>(15));
  return v;
}
// layer l reads its input from xa (even l) / xb (odd l) and leaves its output in the other one
inline bf16* dec_x_in(const DecPart& v, int l) { return (l & 1) ? v.xb : v.xa; }
inline bf16* dec_x_out(const DecPart& v, int l) { return (l & 1) ? v.xa : v.xb; }

int dec_embed(b200_engine* e, const DecPart& v, const int64_t* tokens, int pos, cudaStream_t s) {
  const auto& c = e->cfg;
  B200_REQUIRE(pos >= 0 && pos < e->dec.max_len && pos < c.max_seq_len, "decode: position %d out of range", pos);
  return embed_pe_fwd(tokens + v.r0, e->pf + e->emb, e->pe, v.xa, v.R, 1, c.embed_dim, c.vocab_size,
                      sqrtf(static_cast<float>(c.embed_dim)), s, pos);
}
// self attention block + the cross-attention query:  QKV, attention over the cache (+ append), out-proj + LN1, Q
// seq: [R, max_len] token ids of the hypotheses (columns [0, pos] valid): keys whose token is PAD are masked
int dec_pre(b200_engine* e, const DecPart& v, int l, int cur, int pos, const int64_t* seq, cudaStream_t s) {
  const auto& c = e->cfg;
  auto& d = e->dec;
  const int E = c.embed_dim, H = c.num_heads, hd = E / H;
  const LayerOff& o = e->lo[l];
  const int64_t self_stride = static_cast<int64_t>(d.R) * d.max_len * E;
  bf16* kc_l = d.kcache[cur] + l * self_stride + v.self_off;
  bf16* vc_l = d.vcache[cur] + l * self_stride + v.self_off;
  bf16* x = dec_x_in(v, l);
  RC(linear_dec(x, E, e->ph + o.sa_w, e->pf + o.sa_b, v.qkv, 3 * E, v.R, 3 * E, E, 0, s, d.single_cta));
  RC(attn_decode_append(v.qkv, 3 * E, kc_l, vc_l, d.max_len, pos, v.attn, E, v.R, H, hd, 1.0f / sqrtf(static_cast<float>(hd)), s,
                        seq ? seq + v.r0 * d.max_len : nullptr, d.max_len, c.pad_idx));
  RC(linear_ln_fwd(v.attn, E, e->ph + o.sa_ow, e->pf + o.sa_ob, x, e->pf + o.n1_w, e->pf + o.n1_b, v.pt->parts, d.ksplit_e,
                   v.x1, v.R, E, E, c.ln_eps, s, d.single_cta));
  return linear_dec(v.x1, E, e->ph + o.ca_w, e->pf + o.ca_b, v.qc, E, v.R, E, E, 0, s, d.single_cta);
}
// cross attention over the image's K/V planes of layer l (the HBM-bound stream)
int dec_cross(b200_engine* e, const DecPart& v, int l, cudaStream_t s) {
  const auto& c = e->cfg;
  auto& d = e->dec;
  const int E = c.embed_dim, H = c.num_heads, L = c.num_layers, hd = E / H;
  const int64_t cross_stride = static_cast<int64_t>(d.B) * d.S * E;
  const int ln = (l + 1 < L) ? l + 1 : 0;   // the next launch streams layer l+1 (layer 0 of the next position after the last)
  return attn_decode(v.qc, E, d.kc + l * cross_stride + v.cross_off, d.vc + l * cross_stride + v.cross_off, d.S, d.S, v.attn, E,
                     v.B, d.beam, H, hd, v.mem_pad, 1.0f / sqrtf(static_cast<float>(hd)), s, d.kc + ln * cross_stride + v.cross_off,
                     d.vc + ln * cross_stride + v.cross_off, v.pf_bytes, d.kv_flags, d.attn_fat_grid, v.pt->sched);
}
// cross out-proj + LN2, FFN1, FFN2 + LN3
int dec_post(b200_engine* e, const DecPart& v, int l, cudaStream_t s) {
  const auto& c = e->cfg;
  auto& d = e->dec;
  const int E = c.embed_dim, F = c.ff_dim;
  const LayerOff& o = e->lo[l];
  RC(linear_ln_fwd(v.attn, E, e->ph + o.ca_ow, e->pf + o.ca_ob, v.x1, e->pf + o.n2_w, e->pf + o.n2_b, v.pt->parts, d.ksplit_e,
                   v.x2, v.R, E, E, c.ln_eps, s, d.single_cta));
  RC(linear_dec(v.x2, E, e->ph + o.l1_w, e->pf + o.l1_b, v.h, F, v.R, F, E, c.act, s, d.single_cta));
  return linear_ln_fwd(v.h, F, e->ph + o.l2_w, e->pf + o.l2_b, v.x2, e->pf + o.n3_w, e->pf + o.n3_b, v.pt->parts, d.ksplit_f,
                       dec_x_out(v, l), v.R, E, F, c.ln_eps, s, d.single_cta);
}

// Streams of a (possibly partitioned) generation call: partition p's launch chain runs on S[p] (S[0] is the
// caller's stream), the cross-attention launches of all partitions on SA when the plan asks for it.
// parts_fork / parts_join bracket the work; both are valid eagerly and under stream capture.
struct PartStreams {
  std::vector<cudaStream_t> S;
  cudaStream_t SA = nullptr;
  std::vector<DecPart> view;
};

int parts_fork(b200_engine* e, cudaStream_t s, PartStreams* ps) {
  auto& d = e->dec;
  const int P = static_cast<int>(d.part.size());
  while (static_cast<int>(d.side.size()) < P - 1) {
    cudaStream_t st;
    B200_CHECK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    d.side.push_back(st);
  }
  auto grow = [](std::vector<cudaEvent_t>& v, int n) -> int {
    while (static_cast<int>(v.size()) < n) {
      cudaEvent_t ev;
      B200_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      v.push_back(ev);
    }
    return 0;
  };
  RC(grow(d.ev, P + 1));
  RC(grow(d.evq, P));
  RC(grow(d.eva, P));
  if (d.use_att_stream && !d.att_stream) B200_CHECK_CUDA(cudaStreamCreateWithFlags(&d.att_stream, cudaStreamNonBlocking));
  ps->S.assign(1, s);
  for (int p = 1; p < P; ++p) ps->S.push_back(d.side[p - 1]);
  ps->SA = (d.use_att_stream && P > 1) ? d.att_stream : nullptr;
  ps->view.clear();
  for (int p = 0; p < P; ++p) ps->view.push_back(dec_part(e, d.part[p]));
  if (P > 1) {
    B200_CHECK_CUDA(cudaEventRecord(d.ev[0], s));
    for (int p = 1; p < P; ++p) B200_CHECK_CUDA(cudaStreamWaitEvent(ps->S[p], d.ev[0], 0));
    if (ps->SA) B200_CHECK_CUDA(cudaStreamWaitEvent(ps->SA, d.ev[0], 0));
  }
  return 0;
}
// always called, also after an error, so that a capture is never left forked
void parts_join(b200_engine* e, cudaStream_t s, const PartStreams& ps) {
  auto& d = e->dec;
  const int P = static_cast<int>(ps.S.size());
  for (int p = 1; p < P; ++p) {
    cudaEventRecord(d.ev[p], ps.S[p]);
    cudaStreamWaitEvent(s, d.ev[p], 0);
  }
  if (ps.SA) {
    cudaEventRecord(d.ev[P], ps.SA);
    cudaStreamWaitEvent(s, d.ev[P], 0);
  }
}

// one decode position for every partition: tokens[r] at position pos -> final hidden of partition p in x_out[p]
// (`cur` = live copy of the self-attention cache).  Launch order: per layer, every partition's self-attention
// block, then every partition's cross attention, then every partition's FFN block, so that the attention
// stream (if any) sees the partitions round robin.
int decode_hidden_all(b200_engine* e, const PartStreams& ps, int cur, const int64_t* tokens, int pos, bf16** x_out,
                      const int64_t* seq) {
  auto& d = e->dec;
  const int L = e->cfg.num_layers;
  const int P = static_cast<int>(ps.S.size());
  NvtxRange nvtx_pos("b200.decode.position");
  for (int p = 0; p < P; ++p) RC(dec_embed(e, ps.view[p], tokens, pos, ps.S[p]));
  for (int l = 0; l < L; ++l) {
    for (int p = 0; p < P; ++p) {
      RC(dec_pre(e, ps.view[p], l, cur, pos, seq, ps.S[p]));
      if (ps.SA) B200_CHECK_CUDA(cudaEventRecord(d.evq[p], ps.S[p]));
    }
    for (int p = 0; p < P; ++p) {
      if (ps.SA) {
        B200_CHECK_CUDA(cudaStreamWaitEvent(ps.SA, d.evq[p], 0));
        RC(dec_cross(e, ps.view[p], l, ps.SA));
        B200_CHECK_CUDA(cudaEventRecord(d.eva[p], ps.SA));
      } else {
        RC(dec_cross(e, ps.view[p], l, ps.S[p]));
      }
    }
    for (int p = 0; p < P; ++p) {
      if (ps.SA) B200_CHECK_CUDA(cudaStreamWaitEvent(ps.S[p], d.eva[p], 0));
      RC(dec_post(e, ps.view[p], l, ps.S[p]));
    }
  }
  for (int p = 0; p < P; ++p) x_out[p] = dec_x_in(ps.view[p], L);
  return 0;
}

// fork, run body(streams), join (the join also happens when body fails)
template <class Body>
int with_parts(b200_engine* e, cudaStream_t s, Body body) {
  PartStreams ps;
  RC(parts_fork(e, s, &ps));
  const int rc = body(ps);
  parts_join(e, s, ps);
  return rc;
}

// the plan's tuning knobs, folded into the CUDA-graph cache keys (a sweep re-plans with different settings)
uint64_t dec_tuning_key(const b200_engine::Decode& d) {
  return static_cast<uint64_t>(d.pf_bytes) * 16ull + static_cast<uint64_t>(d.kv_flags) + (static_cast<uint64_t>(d.ksplit_e) << 40) +
         (static_cast<uint64_t>(d.ksplit_f) << 46) + (static_cast<uint64_t>(d.single_cta ? 1 : 0) << 52) +
         (static_cast<uint64_t>(d.attn_fat_grid) << 53) + (static_cast<uint64_t>(d.use_att_stream ? 1 : 0) << 62) + (static_cast<uint64_t>(d.attn_dyn ? 1 : 0) << 61) ^
         (static_cast<uint64_t>(d.gemm_cap) * 0x9E3779B97F4A7C15ull);
}

uint64_t mix_key(std::initializer_list<uint64_t> v) {
  uint64_t h = 1469598103934665603ull;
  for (uint64_t x : v) { h ^= x + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); h *= 1099511628211ull; }
  return h ? h : 1;
}

// Runs `body` (which only enqueues work on `s`) either eagerly or as a cached CUDA graph: the first
// call with a given key runs eagerly (lazy one-time kernel attribute set-up happens there), the
// second captures + instantiates, later calls replay.  A generation step is ~90 small launches, so
// replaying removes the host launch cost that otherwise bounds batch-512 decoding.
template <class Body>
int run_maybe_graphed(b200_engine* e, uint64_t key, cudaStream_t s, Body body) {
  auto& d = e->dec;
  static const bool disabled = getenv("B200_NO_GRAPH") != nullptr;
  if (disabled) return body(s);
  if (d.graph && d.graph_key == key) {
    B200_CHECK_CUDA(cudaGraphLaunch(d.graph, s));
    return 0;
  }
  if (d.seen_key != key) {   // first sighting: eager
    d.seen_key = key;
    return body(s);
  }
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &st);
  if (st != cudaStreamCaptureStatusNone) return body(s);   // caller is already capturing: just enqueue
  if (d.graph) { cudaGraphExecDestroy(d.graph); d.graph = nullptr; d.graph_key = 0; }
  // capture on a private stream (the caller's may be the legacy default stream, which cannot be
  // captured); nothing executes during capture, the instantiated graph is launched on `s`.
  if (!d.cap_stream) B200_CHECK_CUDA(cudaStreamCreateWithFlags(&d.cap_stream, cudaStreamNonBlocking));
  B200_CHECK_CUDA(cudaStreamBeginCapture(d.cap_stream, cudaStreamCaptureModeRelaxed));
  const int rc = body(d.cap_stream);
  cudaGraph_t g = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(d.cap_stream, &g);
  if (rc != 0) { if (g) cudaGraphDestroy(g); return rc; }
  B200_CHECK_CUDA(ce);
  cudaGraphExec_t ex = nullptr;
  const cudaError_t ci = cudaGraphInstantiate(&ex, g, 0);
  cudaGraphDestroy(g);
  B200_CHECK_CUDA(ci);
  d.graph = ex;
  d.graph_key = key;
  B200_CHECK_CUDA(cudaGraphLaunch(d.graph, s));
  return 0;
}

}  // namespace

extern "C" {

int b200_engine_create(const b200_engine_config* cfg, b200_engine** out) {
  B200_REQUIRE(cfg && out, "engine_create: null argument");
  const int E = cfg->embed_dim, H = cfg->num_heads, F = cfg->ff_dim, V = cfg->vocab_size, L = cfg->num_layers;
  B200_REQUIRE(E > 0 && H > 0 && F > 0 && V > 0 && L > 0 && cfg->max_seq_len > 0, "engine_create: non-positive dimension");
  B200_REQUIRE(E % H == 0, "engine_create: embed_dim %d not divisible by num_heads %d", E, H);
  const int hd = E / H;
  B200_REQUIRE(hd == 32 || hd == 64 || hd == 96 || hd == 128, "engine_create: head dim %d not in {32,64,96,128}", hd);
  B200_REQUIRE(E % 64 == 0 && F % 64 == 0 && V % 8 == 0, "engine_create: embed_dim/ff_dim must be multiples of 64 and vocab_size of 8 (got %d/%d/%d)", E, F, V);
  B200_REQUIRE(cfg->enc_dim % 8 == 0, "engine_create: enc_dim %d must be a multiple of 8", cfg->enc_dim);
  b200_engine* e = new b200_engine();
  e->cfg = *cfg;
  if (e->cfg.ln_eps <= 0.f) e->cfg.ln_eps = 1e-5f;
  if (cfg->enc_dim != E) {
    e->proj_w = add_param(e, "projection.weight", static_cast<int64_t>(E) * cfg->enc_dim);
    e->proj_b = add_param(e, "projection.bias", E);
  }
  e->emb = add_param(e, "token_embedding.weight", static_cast<int64_t>(V) * E);
  e->lo.resize(L);
  for (int l = 0; l < L; ++l) {
    const std::string p = "transformer_decoder.layers." + std::to_string(l) + ".";
    LayerOff& o = e->lo[l];
    o.sa_w = add_param(e, p + "self_attn.in_proj_weight", 3LL * E * E);
    o.begin = o.sa_w;
    o.sa_b = add_param(e, p + "self_attn.in_proj_bias", 3LL * E);
    o.sa_ow = add_param(e, p + "self_attn.out_proj.weight", 1LL * E * E);
    o.sa_ob = add_param(e, p + "self_attn.out_proj.bias", E);
    o.ca_w = add_param(e, p + "multihead_attn.in_proj_weight", 3LL * E * E);
    o.ca_b = add_param(e, p + "multihead_attn.in_proj_bias", 3LL * E);
    o.ca_ow = add_param(e, p + "multihead_attn.out_proj.weight", 1LL * E * E);
    o.ca_ob = add_param(e, p + "multihead_attn.out_proj.bias", E);
    o.l1_w = add_param(e, p + "linear1.weight", 1LL * F * E);
    o.l1_b = add_param(e, p + "linear1.bias", F);
    o.l2_w = add_param(e, p + "linear2.weight", 1LL * E * F);
    o.l2_b = add_param(e, p + "linear2.bias", E);
    o.n1_w = add_param(e, p + "norm1.weight", E);
    o.n1_b = add_param(e, p + "norm1.bias", E);
    o.n2_w = add_param(e, p + "norm2.weight", E);
    o.n2_b = add_param(e, p + "norm2.bias", E);
    o.n3_w = add_param(e, p + "norm3.weight", E);
    o.n3_b = add_param(e, p + "norm3.bias", E);
  }
  e->fc_w = add_param(e, "fc_out.weight", static_cast<int64_t>(V) * E);
  e->fc_b = add_param(e, "fc_out.bias", V);
  e->total = (e->total + 63) & ~static_cast<int64_t>(63);
  *out = e;
  return 0;
}

void b200_engine_destroy(b200_engine* e) {
  if (e && e->dec.graph) cudaGraphExecDestroy(e->dec.graph);
  if (e && e->dec.cap_stream) cudaStreamDestroy(e->dec.cap_stream);
  if (e) {
    for (auto st : e->dec.side) cudaStreamDestroy(st);
    for (auto ev : e->dec.ev) cudaEventDestroy(ev);
    for (auto ev : e->dec.evq) cudaEventDestroy(ev);
    for (auto ev : e->dec.eva) cudaEventDestroy(ev);
    if (e->dec.att_stream) cudaStreamDestroy(e->dec.att_stream);
    if (e->bias_stream) cudaStreamDestroy(e->bias_stream);
    if (e->bias_fork) cudaEventDestroy(e->bias_fork);
    if (e->bias_join) cudaEventDestroy(e->bias_join);
    for (auto w : e->wdone) if (w) cudaEventDestroy(w);
  }
  delete e;
}

int64_t b200_engine_param_count(const b200_engine* e) { return e ? e->total : -1; }
int32_t b200_engine_num_params(const b200_engine* e) { return e ? static_cast<int32_t>(e->params.size()) : -1; }

int64_t b200_engine_param_offset(const b200_engine* e, const char* name, int64_t* numel) {
  if (!e || !name) return -1;
  for (const auto& p : e->params)
    if (p.name == name) {
      if (numel) *numel = p.numel;
      return p.offset;
    }
  return -1;
}

int b200_engine_param_name(const b200_engine* e, int32_t index, char* buf, int32_t buflen) {
  B200_REQUIRE(e && buf && index >= 0 && index < static_cast<int32_t>(e->params.size()), "param_name: bad index");
  snprintf(buf, buflen, "%s", e->params[index].name.c_str());
  return 0;
}

int b200_engine_bind(b200_engine* e, float* params_f32, void* params_bf16, float* grads_f32, const float* pe_f32) {
  B200_REQUIRE(e && params_f32 && params_bf16 && pe_f32, "engine_bind: null argument");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(params_f32) & 255) == 0 && (reinterpret_cast<uintptr_t>(params_bf16) & 255) == 0 &&
               (reinterpret_cast<uintptr_t>(grads_f32) & 255) == 0, "engine_bind: arenas must be 256-byte aligned");
  e->pf = params_f32; e->ph = static_cast<bf16*>(params_bf16); e->gf = grads_f32; e->pe = pe_f32;
  // the backward's side stream and its fork / join events are made here, on the device that owns the arenas and
  // outside any stream capture (a first training step may already be captured into a CUDA graph)
  if (grads_f32) {
    cudaPointerAttributes at;
    int dev = -1, cur = -1;
    if (cudaPointerGetAttributes(&at, grads_f32) == cudaSuccess && at.type == cudaMemoryTypeDevice) dev = at.device;
    cudaGetDevice(&cur);
    if (dev >= 0 && dev != cur) cudaSetDevice(dev);
    if (!e->bias_stream) B200_CHECK_CUDA(cudaStreamCreateWithFlags(&e->bias_stream, cudaStreamNonBlocking));
    if (!e->bias_fork) B200_CHECK_CUDA(cudaEventCreateWithFlags(&e->bias_fork, cudaEventDisableTiming));
    if (!e->bias_join) B200_CHECK_CUDA(cudaEventCreateWithFlags(&e->bias_join, cudaEventDisableTiming));
    for (auto& w : e->wdone) if (!w) B200_CHECK_CUDA(cudaEventCreateWithFlags(&w, cudaEventDisableTiming));
    if (dev >= 0 && dev != cur) cudaSetDevice(cur);
  }
  return 0;
}

int b200_engine_set_memory_dtype(b200_engine* e, int32_t is_bf16) {
  B200_REQUIRE(e, "set_memory_dtype: null engine");
  e->mem_bf16 = is_bf16 != 0;
  return 0;
}

int b200_engine_set_dropout(b200_engine* e, float p, uint32_t* state_dev) {
  B200_REQUIRE(e, "set_dropout: null engine");
  B200_REQUIRE(p >= 0.f && p < 1.f, "set_dropout: probability %f outside [0, 1)", p);
  B200_REQUIRE(p == 0.f || (state_dev != nullptr && (reinterpret_cast<uintptr_t>(state_dev) & 7) == 0),
               "set_dropout: p > 0 needs an 8-byte aligned device buffer of two uint32 [seed, counter]");
  e->drop_p = p;
  e->drop_state = state_dev;
  return 0;
}

int64_t b200_engine_workspace_bytes(const b200_engine* e, int32_t B, int32_t T, int32_t S, int32_t mem_dim, int32_t training) {
  if (!e) return -1;
  Plan pl;
  build_plan(e, &pl, nullptr, B, T, S, mem_dim, training);
  return pl.bytes;
}

int b200_engine_set_workspace(b200_engine* e, void* ws, int64_t bytes) {
  B200_REQUIRE(e && ws && (reinterpret_cast<uintptr_t>(ws) & 255) == 0, "set_workspace: null or misaligned workspace");
  e->ws = static_cast<uint8_t*>(ws);
  e->ws_bytes = bytes;
  e->have_saved = false;
  return 0;
}

int b200_engine_forward_logits(b200_engine* e, const int64_t* tokens, const float* memory, const uint8_t* mem_pad,
                               int32_t B, int32_t T, int32_t S, int32_t mem_dim, int32_t training, float* logits, void* stream) {
  B200_REQUIRE(e && tokens && memory && logits, "forward_logits: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RC(run_forward(e, tokens, memory, mem_pad, B, T, S, mem_dim, training, s));
  e->last_tokens = tokens; e->last_targets = nullptr; e->last_mem_pad = mem_pad;
  e->have_saved = training != 0;
  const int E = e->cfg.embed_dim, V = e->cfg.vocab_size;
  GemmProblem g;
  g.M = B * T; g.N = V; g.K = E;
  g.A = e->plan.x_final; g.lda = E; g.B = e->ph + e->fc_w; g.ldb = E;
  g.D = logits; g.ldd = V; g.d_fp32 = true; g.bias = e->pf + e->fc_b; g.split_k = 1;
  return gemm_launch(g, s);
}

static int forward_loss_impl(b200_engine* e, const int64_t* tokens, const int64_t* targets, const float* memory,
                             const uint8_t* mem_pad, int32_t B, int32_t T, int32_t S, int32_t mem_dim,
                             int64_t ignore_index, int32_t training, const int32_t* cu, int32_t rows, float* loss_out,
                             void* stream) {
  B200_REQUIRE(e && tokens && targets && memory && loss_out, "forward_loss: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RC(run_forward(e, tokens, memory, mem_pad, B, T, S, mem_dim, training, s, cu, rows, targets));
  Plan& pl = e->plan;
  if (cu) targets = pl.ptgt;                          // the packed copies (run_forward gathered them)
  e->last_tokens = tokens; e->last_targets = targets; e->last_mem_pad = mem_pad; e->last_ignore = ignore_index;
  e->have_saved = training != 0;
  const int E = e->cfg.embed_dim, V = e->cfg.vocab_size, M = pl.M;
  B200_CHECK_CUDA(cudaMemsetAsync(pl.scalars, 0, 16 * sizeof(float), s));
  RC(b200_lmhead_ce_fwd(pl.x_final, E, e->ph + e->fc_w, E, e->pf + e->fc_b, targets, M, V, E, ignore_index,
                        pl.row_lse, pl.row_loss, pl.scalars, pl.scalars + 1, pl.ce_scratch, stream));
  RC(ce_mean(pl.scalars, pl.scalars + 1, pl.scalars + 4, s));
  B200_CHECK_CUDA(cudaMemcpyAsync(loss_out, pl.scalars + 4, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}

int b200_engine_forward_loss(b200_engine* e, const int64_t* tokens, const int64_t* targets, const float* memory,
                             const uint8_t* mem_pad, int32_t B, int32_t T, int32_t S, int32_t mem_dim,
                             int64_t ignore_index, int32_t training, float* loss_out, void* stream) {
  return forward_loss_impl(e, tokens, targets, memory, mem_pad, B, T, S, mem_dim, ignore_index, training, nullptr, 0, loss_out, stream);
}

int b200_engine_forward_loss_packed(b200_engine* e, const int64_t* tokens, const int64_t* targets, const float* memory,
                                    const uint8_t* mem_pad, int32_t B, int32_t T, int32_t S, int32_t mem_dim,
                                    int64_t ignore_index, int32_t training, const int32_t* cu_seqlens, int32_t total_rows,
                                    float* loss_out, void* stream) {
  B200_REQUIRE(cu_seqlens && total_rows > 0, "forward_loss_packed: cu_seqlens / total_rows missing");
  return forward_loss_impl(e, tokens, targets, memory, mem_pad, B, T, S, mem_dim, ignore_index, training, cu_seqlens, total_rows,
                           loss_out, stream);
}

int b200_engine_backward(b200_engine* e, const float* inv_count_dev, float* dmemory, void* const* bucket_events,
                         int32_t num_bucket_events, void* stream) {
  B200_REQUIRE(e, "backward: null engine");
  B200_REQUIRE(e->have_saved && e->last_targets, "backward: call forward_loss(training=1) first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Plan& pl = e->plan;
  const int E = e->cfg.embed_dim, V = e->cfg.vocab_size, M = pl.M;
  const float* inv = inv_count_dev ? inv_count_dev : pl.scalars + 6;
  RC(b200_lmhead_ce_bwd(pl.x_final, E, e->ph + e->fc_w, E, e->pf + e->fc_b, e->last_targets, M, V, E, e->last_ignore,
                        pl.row_lse, inv, pl.dlogits, V, stream));
  return run_backward(e, false, dmemory, bucket_events, num_bucket_events, s);
}

int b200_engine_backward_parts(b200_engine* e, const float* inv_count_dev, float* dmemory, int32_t first_part,
                               int32_t last_part, void* stream) {
  B200_REQUIRE(e, "backward_parts: null engine");
  B200_REQUIRE(e->have_saved && e->last_targets, "backward_parts: call forward_loss(training=1) first");
  B200_REQUIRE(first_part >= 0 && first_part <= last_part, "backward_parts: bad part range [%d, %d]", first_part, last_part);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Plan& pl = e->plan;
  if (first_part == 0) {
    const int E = e->cfg.embed_dim, V = e->cfg.vocab_size, M = pl.M;
    const float* inv = inv_count_dev ? inv_count_dev : pl.scalars + 6;
    RC(b200_lmhead_ce_bwd(pl.x_final, E, e->ph + e->fc_w, E, e->pf + e->fc_b, e->last_targets, M, V, E, e->last_ignore,
                          pl.row_lse, inv, pl.dlogits, V, stream));
  }
  return run_backward(e, false, dmemory, nullptr, 0, s, first_part, last_part);
}

int b200_engine_backward_from_dlogits(b200_engine* e, const float* dlogits, float* dmemory, void* stream) {
  B200_REQUIRE(e && dlogits, "backward_from_dlogits: null argument");
  B200_REQUIRE(e->have_saved, "backward_from_dlogits: call forward_logits(training=1) first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Plan& pl = e->plan;
  const long long n = static_cast<long long>(pl.B) * pl.T * e->cfg.vocab_size;
  RC(cast_f32_to_bf16(dlogits, pl.dlogits, n, s));
  return run_backward(e, true, dmemory, nullptr, 0, s);
}

int32_t b200_engine_grad_buckets(const b200_engine* e, int64_t* offsets, int64_t* counts, int32_t cap) {
  if (!e) return -1;
  const int L = e->cfg.num_layers;
  const int n = L + 2;
  if (offsets && counts) {
    int i = 0;
    auto put = [&](int64_t off, int64_t end) {
      if (i < cap) { offsets[i] = off; counts[i] = end - off; }
      ++i;
    };
    put(e->fc_w, e->total);
    for (int l = L - 1; l >= 0; --l) put(e->lo[l].begin, (l + 1 < L) ? e->lo[l + 1].begin : e->fc_w);
    put(0, e->lo[0].begin);
  }
  return n;
}

// ---- LM head entry points (also used stand-alone by the tests)
int b200_lmhead_ce_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                       const int64_t* targets, int32_t M, int32_t V, int32_t E, int64_t ignore_index,
                       float* row_lse, float* row_loss, float* loss_sum, float* valid_count, float* scratch,
                       void* stream) {
  B200_REQUIRE(x && w && targets && row_lse && loss_sum && valid_count && scratch, "lmhead_ce_fwd: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n_tiles = gemm_num_n_tiles(V, 256);
  GemmProblem g;
  g.M = M; g.N = V; g.K = E;
  g.A = static_cast<const bf16*>(x); g.lda = ldx; g.B = static_cast<const bf16*>(w); g.ldb = ldw;
  g.bias = bias; g.epi = EPI_CE_FWD; g.targets = targets; g.ignore_index = ignore_index;
  g.part_max = scratch;
  g.part_sum = scratch + static_cast<int64_t>(M) * n_tiles;
  g.tgt_logit = scratch + 2 * static_cast<int64_t>(M) * n_tiles;
  g.split_k = 1;
  RC(gemm_launch(g, s));
  return ce_finalize(g.part_max, g.part_sum, g.tgt_logit, targets, M, n_tiles, ignore_index, row_lse, row_loss,
                     loss_sum, valid_count, s);
}

int b200_lmhead_ce_bwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                       const int64_t* targets, int32_t M, int32_t V, int32_t E, int64_t ignore_index,
                       const float* row_lse, const float* inv_count, void* dlogits, int64_t ldd, void* stream) {
  B200_REQUIRE(x && w && targets && row_lse && inv_count && dlogits, "lmhead_ce_bwd: null argument");
  GemmProblem g;
  g.M = M; g.N = V; g.K = E;
  g.A = static_cast<const bf16*>(x); g.lda = ldx; g.B = static_cast<const bf16*>(w); g.ldb = ldw;
  g.bias = bias; g.epi = EPI_CE_BWD; g.targets = targets; g.ignore_index = ignore_index;
  g.row_lse = row_lse; g.inv_count = inv_count; g.D = dlogits; g.ldd = ldd; g.split_k = 1;
  return gemm_launch(g, static_cast<cudaStream_t>(stream));
}

int b200_lmhead_argmax(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, int32_t M,
                       int32_t V, int32_t E, int64_t* out_ids, float* out_max, float* scratch, void* stream) {
  B200_REQUIRE(x && w && out_ids && scratch, "lmhead_argmax: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n_tiles = gemm_num_n_tiles(V, 128);
  GemmProblem g;
  g.M = M; g.N = V; g.K = E;
  g.A = static_cast<const bf16*>(x); g.lda = ldx; g.B = static_cast<const bf16*>(w); g.ldb = ldw;
  g.bias = bias; g.epi = EPI_ARGMAX;
  g.part_max = scratch; g.part_sum = scratch + static_cast<int64_t>(M) * n_tiles; g.split_k = 1;
  RC(gemm_launch(g, s));
  return argmax_finalize(g.part_max, g.part_sum, M, n_tiles, out_ids, out_max, s);
}


// ---- KV-cached generation ------------------------------------------------------------------
int64_t b200_engine_decode_workspace_bytes(const b200_engine* e, int32_t B, int32_t beam, int32_t S,
                                           int32_t mem_dim, int32_t max_len) {
  if (!e || B <= 0 || beam <= 0 || S <= 0 || max_len <= 0) return -1;
  b200_engine::Decode d;
  build_decode_plan(e, &d, nullptr, B, beam, S, mem_dim, max_len);
  return d.bytes;
}

int b200_engine_decode_begin(b200_engine* e, const float* memory, const uint8_t* mem_pad, int32_t B, int32_t beam,
                             int32_t S, int32_t mem_dim, int32_t max_len, void* ws, int64_t ws_bytes, void* stream) {
  B200_REQUIRE(e && memory && ws, "decode_begin: null argument");
  B200_REQUIRE(e->pf && e->ph, "decode_begin: parameters not bound");
  const auto& c = e->cfg;
  const int E = c.embed_dim, H = c.num_heads, L = c.num_layers, hd = E / H;
  B200_REQUIRE(B > 0 && S > 0 && beam >= 1 && beam <= 4, "decode_begin: B=%d S=%d beam=%d (beam must be 1..4)", B, S, beam);
  B200_REQUIRE(max_len >= 2 && max_len <= c.max_seq_len, "decode_begin: max_len %d outside [2, max_seq_len=%d] (decoder.py:71)", max_len, c.max_seq_len);
  B200_REQUIRE(mem_dim == E || (mem_dim == c.enc_dim && e->proj_w >= 0), "decode_begin: memory width %d matches neither embed_dim nor enc_dim", mem_dim);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "decode_begin: workspace must be 256-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto& d = e->dec;
  d.ready = false;
  build_decode_plan(e, &d, static_cast<uint8_t*>(ws), B, beam, S, mem_dim, max_len);
  B200_REQUIRE(d.bytes <= ws_bytes, "decode_begin: workspace too small (%lld needed, %lld given)", (long long)d.bytes, (long long)ws_bytes);
  d.mem_pad = mem_pad;
  d.cur = 0;
  const int Ms = B * S;
  if (e->mem_bf16) {
    B200_REQUIRE((reinterpret_cast<uintptr_t>(memory) & 15) == 0, "decode_begin: bf16 memory must be 16-byte aligned");
    const bool same = d.memp == d.mem16;
    d.mem16 = reinterpret_cast<bf16*>(const_cast<float*>(memory));     // only read inside this call
    if (same) d.memp = d.mem16;
  } else {
    RC(cast_f32_to_bf16(memory, d.mem16, static_cast<long long>(Ms) * mem_dim, s));
  }
  if (mem_dim != E)
    RC(linear_fwd(d.mem16, mem_dim, e->ph + e->proj_w, e->pf + e->proj_b, d.memp, E, Ms, E, mem_dim, 0, nullptr, 0, s));
  // image-side K/V: projected ONCE per image and layer (the reference redoes this every token)
  for (int l = 0; l < L; ++l) {
    const LayerOff& o = e->lo[l];
    RC(linear_fwd(d.memp, E, e->ph + o.ca_w + static_cast<int64_t>(E) * E, e->pf + o.ca_b + E, d.kv_tmp, 2 * E, Ms, 2 * E, E, 0, nullptr, 0, s));
    RC(kv_to_head_major(d.kv_tmp, d.kc + static_cast<int64_t>(l) * Ms * E, d.vc + static_cast<int64_t>(l) * Ms * E, B, S, H, hd, s));
  }
  B200_CHECK_CUDA(cudaMemsetAsync(d.sched, 0, 16 * sizeof(int), s));
  d.ready = true;
  return 0;
}

// LM head of one partition: greedy ids (argmax fused into the GEMM epilogue) from its final hidden state x
static int decode_argmax_part(b200_engine* e, const DecPart& v, const bf16* x, int64_t* next_ids, cudaStream_t s) {
  auto& d = e->dec;
  const int E = e->cfg.embed_dim, V = e->cfg.vocab_size;
  const int64_t n_tiles = gemm_num_n_tiles(V, 128);
  return b200_lmhead_argmax(x, E, e->ph + e->fc_w, E, e->pf + e->fc_b, v.R, V, E, next_ids + v.r0, d.best + v.r0,
                            d.scratch + 2 * v.r0 * n_tiles, s);
}

int b200_engine_decode_plan_info(const b200_engine* e, int32_t* info, int32_t n) {
  B200_REQUIRE(e && info && n >= 0, "decode_plan_info: null argument");
  B200_REQUIRE(e->dec.ready, "decode_plan_info: call decode_begin first");
  const auto& d = e->dec;
  const int32_t v[6] = {static_cast<int32_t>(d.part.size()), d.attn_fat_grid, d.ksplit_e, d.ksplit_f, d.gemm_cap, d.kv_flags};
  for (int i = 0; i < n && i < 6; ++i) info[i] = v[i];
  return 0;
}

int b200_engine_decode_step(b200_engine* e, const int64_t* tokens_in, int32_t pos, int64_t* next_ids, void* stream) {
  B200_REQUIRE(e && tokens_in && next_ids, "decode_step: null argument");
  B200_REQUIRE(e->dec.ready, "decode_step: call decode_begin first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GemmGridCap grid_cap(e->dec.gemm_cap, (e->dec.kv_flags & 4) != 0);
  return with_parts(e, s, [&](const PartStreams& ps) -> int {
    std::vector<bf16*> x(ps.S.size(), nullptr);
    // the caller feeds one column at a time: keep the prefix (for the PAD-key mask) in the plan's sequence buffer
    for (size_t p = 0; p < ps.S.size(); ++p) {
      const DecPart& v = ps.view[p];
      RC(store_col_i64(e->dec.seq[0] + v.r0 * e->dec.max_len, tokens_in + v.r0, v.R, e->dec.max_len, pos, ps.S[p]));
    }
    RC(decode_hidden_all(e, ps, e->dec.cur, tokens_in, pos, x.data(), e->dec.seq[0]));
    for (size_t p = 0; p < ps.S.size(); ++p) RC(decode_argmax_part(e, ps.view[p], x[p], next_ids, ps.S[p]));
    return 0;
  });
}

int b200_engine_generate_greedy(b200_engine* e, int64_t start_id, int64_t end_id, int32_t max_len,
                                int32_t stop_check_interval, int64_t* out_tokens, int32_t* out_len, void* stream) {
  B200_REQUIRE(e && out_tokens && out_len, "generate_greedy: null argument");
  auto& d = e->dec;
  B200_REQUIRE(d.ready && d.beam == 1, "generate_greedy: call decode_begin(beam=1) first");
  B200_REQUIRE(max_len >= 2 && max_len <= d.max_len, "generate_greedy: max_len %d exceeds the decode plan's %d", max_len, d.max_len);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GemmGridCap grid_cap(d.gemm_cap, (d.kv_flags & 4) != 0);
  const int R = d.R;
  const long long pad = e->cfg.pad_idx;
  int64_t* toks = d.seq[0];          // [R, d.max_len] staging inside the workspace (stable address for the graph)
  const int ld = d.max_len;
  auto prologue = [&](cudaStream_t ws) -> int {
    RC(fill_i64(toks, static_cast<long long>(R) * ld, pad, ws));
    RC(fill_col_i64(toks, R, ld, start_id, ws));
    RC(fill_i64(d.cur_tok, R, start_id, ws));
    B200_CHECK_CUDA(cudaMemsetAsync(d.fin[0], 0, R, ws));
    B200_CHECK_CUDA(cudaMemsetAsync(d.n_finished, 0, sizeof(int), ws));
    B200_CHECK_CUDA(cudaMemsetAsync(d.out_len, 0, sizeof(int) * R, ws));
    return 0;
  };
  // positions [p0, p1): every partition's chain on its own stream (rows are independent)
  auto steps = [&](const PartStreams& ps, int p0, int p1) -> int {
    std::vector<bf16*> x(ps.S.size(), nullptr);
    for (int pos = p0; pos < p1; ++pos) {
      RC(decode_hidden_all(e, ps, 0, d.cur_tok, pos, x.data(), toks));
      for (size_t p = 0; p < ps.S.size(); ++p) {
        const DecPart& v = ps.view[p];
        RC(decode_argmax_part(e, v, x[p], d.ids, ps.S[p]));
        RC(greedy_update(d.ids + v.r0, d.cur_tok + v.r0, toks + v.r0 * ld, d.out_len + v.r0, d.fin[0] + v.r0, d.n_finished,
                         v.B, ld, pos, end_id, pad, ps.S[p]));
      }
    }
    return 0;
  };
  if (stop_check_interval > 0) {
    RC(prologue(s));
    for (int p0 = 0; p0 + 1 < max_len; p0 += stop_check_interval) {
      const int p1 = (p0 + stop_check_interval < max_len - 1) ? p0 + stop_check_interval : max_len - 1;
      RC(with_parts(e, s, [&](const PartStreams& ps) -> int { return steps(ps, p0, p1); }));
      int nf = 0;   // the reference's early exit (model.py:239-240), batched
      B200_CHECK_CUDA(cudaMemcpyAsync(&nf, d.n_finished, sizeof(int), cudaMemcpyDeviceToHost, s));
      B200_CHECK_CUDA(cudaStreamSynchronize(s));
      if (nf >= R) break;
    }
  } else {
    const uint64_t key = mix_key({reinterpret_cast<uint64_t>(d.kc), static_cast<uint64_t>(R), static_cast<uint64_t>(d.S),
                                  static_cast<uint64_t>(d.max_len), static_cast<uint64_t>(max_len), static_cast<uint64_t>(start_id),
                                  static_cast<uint64_t>(end_id), reinterpret_cast<uint64_t>(d.mem_pad),
                                  dec_tuning_key(d), static_cast<uint64_t>(d.part.size()), 1ull});
    RC(run_maybe_graphed(e, key, s, [&](cudaStream_t ws) -> int {
      RC(prologue(ws));
      return with_parts(e, ws, [&](const PartStreams& ps) -> int { return steps(ps, 0, max_len - 1); });
    }));
  }
  B200_CHECK_CUDA(cudaMemcpy2DAsync(out_tokens, static_cast<size_t>(max_len) * sizeof(int64_t), toks,
                                    static_cast<size_t>(ld) * sizeof(int64_t), static_cast<size_t>(max_len) * sizeof(int64_t),
                                    R, cudaMemcpyDeviceToDevice, s));
  B200_CHECK_CUDA(cudaMemcpyAsync(out_len, d.out_len, sizeof(int) * R, cudaMemcpyDeviceToDevice, s));
  return 0;
}

int b200_engine_generate_beam(b200_engine* e, int64_t start_id, int64_t end_id, int32_t max_len, int64_t* out_tokens,
                              int32_t* out_len, float* out_score, void* stream) {
  B200_REQUIRE(e && out_tokens && out_len, "generate_beam: null argument");
  auto& d = e->dec;
  B200_REQUIRE(d.ready && d.beam >= 1, "generate_beam: call decode_begin first");
  B200_REQUIRE(d.beam == 1 || d.logits != nullptr, "generate_beam: decode plan has no logits buffer");
  B200_REQUIRE(max_len >= 2 && max_len <= d.max_len, "generate_beam: max_len %d exceeds the decode plan's %d", max_len, d.max_len);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GemmGridCap grid_cap(d.gemm_cap, (d.kv_flags & 4) != 0);
  const auto& c = e->cfg;
  const int E = c.embed_dim, V = c.vocab_size, H = c.num_heads, L = c.num_layers, hd = E / H;
  const int R = d.R, beam = d.beam;
  B200_REQUIRE(d.logits != nullptr, "generate_beam: decode_begin was called with beam=1; use generate_greedy");
  const uint64_t key = mix_key({reinterpret_cast<uint64_t>(d.kc), static_cast<uint64_t>(R), static_cast<uint64_t>(d.S),
                                static_cast<uint64_t>(d.max_len), static_cast<uint64_t>(max_len), static_cast<uint64_t>(start_id),
                                static_cast<uint64_t>(end_id), reinterpret_cast<uint64_t>(d.mem_pad), static_cast<uint64_t>(beam),
                                reinterpret_cast<uint64_t>(out_tokens), reinterpret_cast<uint64_t>(out_len),
                                reinterpret_cast<uint64_t>(out_score), dec_tuning_key(d),
                                static_cast<uint64_t>(d.part.size()), 2ull});
  // every partition (a range of images with all their hypotheses) runs the whole search on its stream;
  // the live copies of the cache / sequences / scores alternate identically in every partition
  RC(run_maybe_graphed(e, key, s, [&](cudaStream_t ws) -> int {
    RC(fill_i64(d.seq[0], static_cast<long long>(R) * d.max_len, c.pad_idx, ws));
    RC(fill_col_i64(d.seq[0], R, d.max_len, start_id, ws));
    RC(fill_i64(d.cur_tok, R, start_id, ws));
    B200_CHECK_CUDA(cudaMemsetAsync(d.fin[0], 0, R, ws));
    B200_CHECK_CUDA(cudaMemsetAsync(d.scores[0], 0, sizeof(float) * R, ws));
    return with_parts(e, ws, [&](const PartStreams& ps) -> int {
      const size_t P = ps.S.size();
      std::vector<bf16*> x(P, nullptr);
      int cur = 0, cs = 0, n_tok = 1;
      for (int pos = 0; pos + 1 < max_len; ++pos) {
        RC(decode_hidden_all(e, ps, cur, d.cur_tok, pos, x.data(), d.seq[cs]));
        for (size_t p = 0; p < P; ++p) {
          const DecPart& v = ps.view[p];
          const int64_t r0 = v.r0;
          cudaStream_t st = ps.S[p];
          float* logits = d.logits + r0 * V;
          GemmProblem g;
          g.M = v.R; g.N = V; g.K = E;
          g.A = x[p]; g.lda = E; g.B = e->ph + e->fc_w; g.ldb = E;
          g.D = logits; g.ldd = V; g.d_fp32 = true; g.bias = e->pf + e->fc_b; g.split_k = 1;
          RC(gemm_launch(g, st));
          RC(beam_topk(logits, d.scores[cs] + r0, d.fin[cs] + r0, v.B, beam, V, end_id, pos == 0, d.cur_tok + r0, d.parent + r0,
                       d.scores[cs ^ 1] + r0, st));
          RC(beam_advance(d.seq[cs] + r0 * d.max_len, d.seq[cs ^ 1] + r0 * d.max_len, d.fin[cs] + r0, d.fin[cs ^ 1] + r0,
                          d.cur_tok + r0, d.parent + r0, v.R, beam, d.max_len, pos, end_id, st));
          // the self-attention cache follows the surviving hypotheses
          for (int l = 0; l < L; ++l) {
            const int64_t off = (static_cast<int64_t>(l) * R + r0) * d.max_len * E;
            RC(cache_reorder(d.kcache[cur] + off, d.vcache[cur] + off, d.kcache[cur ^ 1] + off, d.vcache[cur ^ 1] + off,
                             d.parent + r0, v.B, beam, H, hd, d.max_len, pos, st));
          }
        }
        cur ^= 1;
        cs ^= 1;
        ++n_tok;
      }
      for (size_t p = 0; p < P; ++p) {
        const DecPart& v = ps.view[p];
        const int b0 = v.pt->b0;
        RC(beam_finalize(d.seq[cs] + v.r0 * d.max_len, d.scores[cs] + v.r0, v.B, beam, d.max_len, n_tok, end_id, c.pad_idx,
                         d.out_tok + static_cast<int64_t>(b0) * d.max_len, d.out_len + b0, d.out_score + b0, ps.S[p]));
      }
      return 0;
    });
  }));
  const int B = d.B;
  B200_CHECK_CUDA(cudaMemcpyAsync(out_tokens, d.out_tok, sizeof(int64_t) * B * d.max_len, cudaMemcpyDeviceToDevice, s));
  B200_CHECK_CUDA(cudaMemcpyAsync(out_len, d.out_len, sizeof(int) * B, cudaMemcpyDeviceToDevice, s));
  if (out_score) B200_CHECK_CUDA(cudaMemcpyAsync(out_score, d.out_score, sizeof(float) * B, cudaMemcpyDeviceToDevice, s));
  return 0;
}

}  // extern "C"
