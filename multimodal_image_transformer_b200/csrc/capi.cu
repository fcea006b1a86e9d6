// extern "C" surface declared in include/b200_decoder.h: argument validation + dispatch only.
#include "../../include/b200_decoder.h"
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "attention.cuh"
#include "attention_tc.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <vector>

namespace b200 {
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = getenv("B200_NO_PDL") == nullptr;
  return on;
}

static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct GemmProf {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  std::vector<double> flops;
  size_t used = 0;       // events handed out (2 per launch)
  size_t cap = 0;
};
static GemmProf g_prof;
bool gemm_profile_enabled() { return g_prof.on && g_prof.used + 2 <= g_prof.cap; }
void gemm_profile_record(cudaStream_t s, bool begin, double flops) {
  if (!g_prof.on || g_prof.used >= g_prof.cap) return;
  cudaEventRecord(g_prof.pool[g_prof.used++], s);
  if (begin) g_prof.flops.push_back(flops);
}
}  // namespace b200

using namespace b200;

extern "C" {

int b200_version(void) { return B200_ABI_VERSION; }
const char* b200_last_error(void) { return g_err; }

long long b200_launch_count(void) { return g_launches.load(); }

int b200_gemm_profile_begin(int32_t max_launches) {
  B200_REQUIRE(max_launches > 0 && max_launches <= (1 << 20), "gemm_profile_begin: bad capacity");
  const size_t need = static_cast<size_t>(max_launches) * 2;
  while (g_prof.pool.size() < need) {
    cudaEvent_t ev;
    B200_CHECK_CUDA(cudaEventCreate(&ev));
    g_prof.pool.push_back(ev);
  }
  g_prof.cap = need;
  g_prof.used = 0;
  g_prof.flops.clear();
  g_prof.on = true;
  return 0;
}

int b200_gemm_profile_end(int32_t* n_launches, double* total_ms, double* total_flops, float* per_launch_ms,
                          double* per_launch_flops, int32_t cap) {
  g_prof.on = false;
  const size_t n = g_prof.used / 2;
  double ms_sum = 0.0, fl_sum = 0.0;
  for (size_t i = 0; i < n; ++i) {
    float ms = 0.f;
    B200_CHECK_CUDA(cudaEventElapsedTime(&ms, g_prof.pool[2 * i], g_prof.pool[2 * i + 1]));
    ms_sum += ms;
    fl_sum += g_prof.flops[i];
    if (per_launch_ms && static_cast<int32_t>(i) < cap) per_launch_ms[i] = ms;
    if (per_launch_flops && static_cast<int32_t>(i) < cap) per_launch_flops[i] = g_prof.flops[i];
  }
  if (n_launches) *n_launches = static_cast<int32_t>(n);
  if (total_ms) *total_ms = ms_sum;
  if (total_flops) *total_flops = fl_sum;
  return 0;
}

int b200_check_device(int dev) {
  int major = 0, minor = 0;
  B200_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  B200_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  B200_REQUIRE(major == 10 && minor == 0, "device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
  return 0;
}

static int to_problem(const b200_gemm_args* a, GemmProblem* p) {
  B200_REQUIRE(a != nullptr, "gemm: null args");
  p->M = a->M; p->N = a->N; p->K = a->K;
  p->A = static_cast<const bf16*>(a->A); p->lda = a->lda; p->a_mn = a->a_mn_major != 0;
  p->B = static_cast<const bf16*>(a->B); p->ldb = a->ldb; p->b_mn = a->b_mn_major != 0;
  p->D = a->D; p->ldd = a->ldd; p->d_fp32 = a->d_fp32 != 0; p->accumulate = a->accumulate != 0;
  p->bias = a->bias;
  p->residual = static_cast<const bf16*>(a->residual); p->ldr = a->ldr;
  p->relu_mask = static_cast<const bf16*>(a->relu_mask); p->ldm = a->ldm;
  p->act = a->act; p->split_k = a->split_k; p->block_n = a->block_n;
  return 0;
}

int b200_gemm(const b200_gemm_args* a, void* stream) {
  GemmProblem p;
  if (int rc = to_problem(a, &p)) return rc;
  return gemm_launch(p, static_cast<cudaStream_t>(stream));
}

int b200_gemm_check(const b200_gemm_args* a, void* stream) {
  GemmProblem p;
  if (int rc = to_problem(a, &p)) return rc;
  return gemm_check_launch(p, static_cast<cudaStream_t>(stream));
}

#define S_(x) static_cast<cudaStream_t>(x)
#define BF(x) static_cast<bf16*>(x)
#define CBF(x) static_cast<const bf16*>(x)

int b200_embed_pe_fwd(const int64_t* tokens, const float* emb, const float* pe, void* x, int32_t B, int32_t T,
                      int32_t E, int32_t V, float scale, void* stream) {
  B200_REQUIRE(tokens && emb && pe && x, "embed_pe_fwd: null argument");
  return embed_pe_fwd(tokens, emb, pe, BF(x), B, T, E, V, scale, S_(stream));
}
int b200_embed_bwd(const int64_t* tokens, const void* dx, float* demb, int32_t B, int32_t T, int32_t E, int32_t V,
                   int64_t pad_idx, float scale, void* stream) {
  B200_REQUIRE(tokens && dx && demb, "embed_bwd: null argument");
  return embed_bwd(tokens, CBF(dx), demb, B, T, E, V, pad_idx, scale, S_(stream));
}
int b200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                       int32_t rows, int32_t E, float eps, void* stream) {
  B200_REQUIRE(x && gamma && beta && y, "layernorm_fwd: null argument");
  return layernorm_fwd(CBF(x), gamma, beta, BF(y), mean, rstd, rows, E, eps, S_(stream));
}
int b200_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                       void* dx, float* dgamma, float* dbeta, float* dxsum, int32_t rows, int32_t E, void* stream) {
  B200_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, "layernorm_bwd: null argument");
  return layernorm_bwd(CBF(dy), CBF(x), gamma, mean, rstd, BF(dx), dgamma, dbeta, dxsum, rows, E, S_(stream));
}
int b200_colsum(const void* x, int64_t ldx, float* out, int32_t M, int32_t N, void* stream) {
  B200_REQUIRE(x && out, "colsum: null argument");
  return colsum(CBF(x), ldx, out, M, N, S_(stream));
}
int b200_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  B200_REQUIRE((src && dst) || n == 0, "cast: null argument");
  return cast_f32_to_bf16(src, BF(dst), n, S_(stream));
}
int b200_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream) {
  B200_REQUIRE((src && dst) || n == 0, "cast: null argument");
  return cast_bf16_to_f32(CBF(src), dst, n, S_(stream));
}

static void to_attn(const b200_attn_fwd_args* a, AttnArgs* o) {
  o->q = CBF(a->q); o->q_bs = a->q_bs; o->q_ts = a->q_ts;
  o->k = CBF(a->k); o->k_bs = a->k_bs; o->k_ts = a->k_ts;
  o->v = CBF(a->v); o->v_bs = a->v_bs; o->v_ts = a->v_ts;
  o->o = BF(a->o); o->o_bs = a->o_bs; o->o_ts = a->o_ts;
  o->lse = a->lse; o->B = a->B; o->H = a->H; o->Tq = a->Tq; o->Tk = a->Tk; o->hd = a->hd;
  o->causal = a->causal; o->key_tokens = a->key_tokens; o->pad_idx = a->pad_idx;
  o->key_pad_mask = a->key_pad_mask; o->scale = a->scale;
  o->cu_q = a->cu_q; o->cu_k = a->cu_k; o->total_q = a->total_q; o->total_k = a->total_k;
}
int b200_attn_fwd(const b200_attn_fwd_args* a, void* stream) {
  B200_REQUIRE(a, "attn_fwd: null args");
  AttnArgs o;
  to_attn(a, &o);
  return attn_fwd(o, S_(stream));
}
int b200_attn_bwd(const b200_attn_bwd_args* a, void* stream) {
  B200_REQUIRE(a, "attn_bwd: null args");
  AttnArgs o;
  to_attn(&a->f, &o);
  AttnGrads g;
  g.d_o = CBF(a->d_o); g.do_bs = a->do_bs; g.do_ts = a->do_ts;
  g.dq = BF(a->dq); g.dq_bs = a->dq_bs; g.dq_ts = a->dq_ts;
  g.dk = BF(a->dk); g.dk_bs = a->dk_bs; g.dk_ts = a->dk_ts;
  g.dv = BF(a->dv); g.dv_bs = a->dv_bs; g.dv_ts = a->dv_ts;
  return attn_bwd(o, g, S_(stream));
}

int b200_attn_tc_trace(void* stamps_dev) {
  attn_tc_set_trace(static_cast<long long*>(stamps_dev));
  return 0;
}

int b200_grad_sumsq(const float* grad, int64_t n, float* sumsq, void* stream) {
  B200_REQUIRE(grad && sumsq, "grad_sumsq: null argument");
  return grad_sumsq(grad, n, sumsq, S_(stream));
}
int b200_adamw_step(float* param, void* param_bf16, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                    const float* sumsq, float max_norm, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int32_t step, void* stream) {
  B200_REQUIRE(param && grad && exp_avg && exp_avg_sq, "adamw_step: null argument");
  return adamw_step(param, BF(param_bf16), grad, exp_avg, exp_avg_sq, n, sumsq, max_norm, lr, beta1, beta2, eps,
                    weight_decay, step, S_(stream));
}
int b200_adamw_step_dev(float* param, void* param_bf16, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                        const float* sumsq, float max_norm, const float* lr_dev, float beta1, float beta2, float eps,
                        float weight_decay, int32_t* step_dev, void* stream) {
  B200_REQUIRE(param && grad && exp_avg && exp_avg_sq && lr_dev && step_dev, "adamw_step_dev: null argument");
  return adamw_step(param, BF(param_bf16), grad, exp_avg, exp_avg_sq, n, sumsq, max_norm, 0.f, beta1, beta2, eps,
                    weight_decay, 0, S_(stream), step_dev, lr_dev);
}

}  // extern "C"
