// Memory-bound kernels of KV-cached generation (replaces the O(T^2) full-prefix recompute of the
// reference greedy loop, model.py:221-240, where every step re-runs the whole decoder and
// re-projects the image memory in every layer).  Per step and layer the dominant traffic is the
// per-image cross-attention K/V (S x E x 2 bf16), streamed exactly once with coalesced 128-bit
// loads from a head-major layout; all beams of an image share that stream.
#include "decode.cuh"
#include <math.h>

namespace b200 {

static inline int cdiv(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------
__global__ void kv_to_head_major_kernel(const bf16* __restrict__ kv, bf16* __restrict__ k_hm,
                                        bf16* __restrict__ v_hm, int B, int S, int H, int hd) {
  const int vpr = hd >> 3;  // 16-byte vectors per head row
  const long long total = static_cast<long long>(B) * S * H * vpr;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % vpr);
  const int h = static_cast<int>((idx / vpr) % H);
  const int s = static_cast<int>((idx / (static_cast<long long>(vpr) * H)) % S);
  const int b = static_cast<int>(idx / (static_cast<long long>(vpr) * H * S));
  const int E = H * hd;
  const bf16* src = kv + (static_cast<long long>(b) * S + s) * 2 * E + h * hd + c * 8;
  const long long dst = ((static_cast<long long>(b) * H + h) * S + s) * hd + c * 8;
  *reinterpret_cast<uint4*>(k_hm + dst) = ldg_nc_v4(src);
  *reinterpret_cast<uint4*>(v_hm + dst) = ldg_nc_v4(src + E);
}

int kv_to_head_major(const bf16* kv, bf16* k_hm, bf16* v_hm, int B, int S, int H, int hd, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * S * H * (hd / 8);
  kv_to_head_major_kernel<<<cdiv(total, 256), 256, 0, s>>>(kv, k_hm, v_hm, B, S, H, hd);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void kv_append_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ kcache,
                                 bf16* __restrict__ vcache, int R, int H, int hd, int max_len, int pos) {
  const int vpr = hd >> 3;
  const long long total = static_cast<long long>(R) * H * vpr;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % vpr);
  const int h = static_cast<int>((idx / vpr) % H);
  const int r = static_cast<int>(idx / (static_cast<long long>(vpr) * H));
  const int E = H * hd;
  const bf16* src = qkv + static_cast<long long>(r) * 3 * E + E + h * hd + c * 8;
  const long long dst = ((static_cast<long long>(r) * H + h) * max_len + pos) * hd + c * 8;
  *reinterpret_cast<uint4*>(kcache + dst) = *reinterpret_cast<const uint4*>(src);
  *reinterpret_cast<uint4*>(vcache + dst) = *reinterpret_cast<const uint4*>(src + E);
}

int kv_append(const bf16* qkv, bf16* kcache, bf16* vcache, int R, int H, int hd, int max_len, int pos, cudaStream_t s) {
  const long long total = static_cast<long long>(R) * H * (hd / 8);
  kv_append_kernel<<<cdiv(total, 256), 256, 0, s>>>(qkv, kcache, vcache, R, H, hd, max_len, pos);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// single-query attention over a resident-in-HBM K/V stream
// ------------------------------------------------------------------------------------------
template <int HD, int NQ>
__global__ void __launch_bounds__(128)
attn_decode_kernel(const bf16* __restrict__ q, long long q_rs, const bf16* __restrict__ k,
                   const bf16* __restrict__ v, int kv_len, int nkeys, bf16* __restrict__ o, long long o_rs,
                   int nq, int H, const unsigned char* __restrict__ key_pad, float scale) {
  constexpr int GL = (HD <= 32) ? 4 : (HD <= 64 ? 8 : 16);   // lanes cooperating on one key row
  constexpr int KPW = 32 / GL;                               // keys per warp per iteration
  constexpr int KPI = 4 * KPW;                               // keys per CTA per iteration
  extern __shared__ float sm_dec[];
  float* sc = sm_dec;                      // [NQ][nkeys]
  float* red = sm_dec + NQ * nkeys;        // [4][NQ][HD]
  __shared__ float s_inv[NQ];

  const int h = blockIdx.x, g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gl = lane % GL, kg = lane / GL;
  const int d0 = gl * 8;
  const bool active = d0 < HD;
  const long long kv_base = (static_cast<long long>(g) * H + h) * kv_len * HD;
  const bf16* kp = k + kv_base;
  const bf16* vp = v + kv_base;

  float qr[NQ][8];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) qr[i][j] = 0.f;
    if (active && i < nq) {
      const uint4 u = *reinterpret_cast<const uint4*>(q + (static_cast<long long>(g) * nq + i) * q_rs + h * HD + d0);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      qr[i][0] = a.x; qr[i][1] = a.y; qr[i][2] = b.x; qr[i][3] = b.y;
      qr[i][4] = c.x; qr[i][5] = c.y; qr[i][6] = d.x; qr[i][7] = d.y;
    }
  }

  // ---- phase 1: scores
#pragma unroll 4
  for (int j0 = 0; j0 < nkeys; j0 += KPI) {
    const int j = j0 + warp * KPW + kg;
    float part[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) part[i] = 0.f;
    if (active && j < nkeys) {
      const uint4 u = ldg_nc_v4(kp + static_cast<long long>(j) * HD + d0);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      const float kf[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
#pragma unroll
        for (int e = 0; e < 8; ++e) part[i] = fmaf(qr[i][e], kf[e], part[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
#pragma unroll
      for (int off = GL / 2; off > 0; off >>= 1) part[i] += __shfl_xor_sync(0xffffffffu, part[i], off);
    }
    if (gl == 0 && j < nkeys) {
      const bool masked = key_pad != nullptr && key_pad[static_cast<long long>(g) * nkeys + j] != 0;
#pragma unroll
      for (int i = 0; i < NQ; ++i) sc[i * nkeys + j] = masked ? -INFINITY : part[i] * scale;
    }
  }
  __syncthreads();

  // ---- softmax: warp i normalises query i
  if (warp < NQ) {
    float* row = sc + warp * nkeys;
    float m = -INFINITY;
    for (int j = lane; j < nkeys; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < nkeys; j += 32) {
      const float p = (m == -INFINITY) ? 0.f : __expf(row[j] - m);
      row[j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    if (lane == 0) s_inv[warp] = sum > 0.f ? 1.f / sum : 0.f;
  }
  __syncthreads();

  // ---- phase 2: o = P V
  float acc[NQ][8];
#pragma unroll
  for (int i = 0; i < NQ; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;
#pragma unroll 4
  for (int j0 = 0; j0 < nkeys; j0 += KPI) {
    const int j = j0 + warp * KPW + kg;
    if (active && j < nkeys) {
      const uint4 u = ldg_nc_v4(vp + static_cast<long long>(j) * HD + d0);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      const float vf[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const float p = sc[i * nkeys + j];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(p, vf[e], acc[i][e]);
      }
    }
  }
  // combine the KPW key groups of a warp, then the 4 warps
#pragma unroll
  for (int i = 0; i < NQ; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int off = GL; off < 32; off <<= 1) acc[i][e] += __shfl_xor_sync(0xffffffffu, acc[i][e], off);
  if (kg == 0 && active) {
#pragma unroll
    for (int i = 0; i < NQ; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) red[(warp * NQ + i) * HD + d0 + e] = acc[i][e];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < NQ * HD; idx += blockDim.x) {
    const int i = idx / HD, d = idx % HD;
    if (i < nq) {
      const float val = (red[(0 * NQ + i) * HD + d] + red[(1 * NQ + i) * HD + d] + red[(2 * NQ + i) * HD + d] +
                         red[(3 * NQ + i) * HD + d]) * s_inv[i];
      o[(static_cast<long long>(g) * nq + i) * o_rs + h * HD + d] = __float2bfloat16(val);
    }
  }
}

template <int HD, int NQ>
static int launch_attn_decode(const bf16* q, long long q_rs, const bf16* k, const bf16* v, int kv_len, int nkeys,
                              bf16* o, long long o_rs, int groups, int nq, int H, const unsigned char* key_pad,
                              float scale, cudaStream_t s) {
  const size_t smem = (static_cast<size_t>(NQ) * nkeys + 4 * NQ * HD) * sizeof(float);
  attn_decode_kernel<HD, NQ><<<dim3(H, groups), 128, smem, s>>>(q, q_rs, k, v, kv_len, nkeys, o, o_rs, nq, H, key_pad, scale);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int attn_decode(const bf16* q, long long q_rs, const bf16* k, const bf16* v, int kv_len, int nkeys, bf16* o,
                long long o_rs, int groups, int nq, int H, int hd, const unsigned char* key_pad, float scale,
                cudaStream_t s) {
  B200_REQUIRE(nq >= 1 && nq <= 4, "attn_decode: %d queries per group (supported: 1..4)", nq);
  B200_REQUIRE(nkeys >= 1 && nkeys <= 1024, "attn_decode: nkeys %d out of range", nkeys);
#define B200_AD(HDV)                                                                                         \
  do {                                                                                                       \
    if (nq == 1) return launch_attn_decode<HDV, 1>(q, q_rs, k, v, kv_len, nkeys, o, o_rs, groups, nq, H, key_pad, scale, s); \
    return launch_attn_decode<HDV, 4>(q, q_rs, k, v, kv_len, nkeys, o, o_rs, groups, nq, H, key_pad, scale, s); \
  } while (0)
  switch (hd) {
    case 32: B200_AD(32);
    case 64: B200_AD(64);
    case 96: B200_AD(96);
    case 128: B200_AD(128);
    default: B200_REQUIRE(false, "attn_decode: head dim %d not in {32,64,96,128}", hd);
  }
#undef B200_AD
}

// ------------------------------------------------------------------------------------------
__global__ void greedy_update_kernel(const int64_t* __restrict__ next_ids, int64_t* __restrict__ cur,
                                     int64_t* __restrict__ out_tokens, int* __restrict__ out_len,
                                     unsigned char* __restrict__ finished, int* __restrict__ n_finished, int R,
                                     int max_len, int pos, long long end_id, long long pad_id) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  long long tok = next_ids[r];
  if (finished[r]) {
    tok = pad_id;
  } else {
    out_len[r] = pos + 2;   // START + (pos+1) generated tokens
    if (tok == end_id) {
      finished[r] = 1;
      atomicAdd(n_finished, 1);
    }
  }
  out_tokens[static_cast<long long>(r) * max_len + pos + 1] = tok;
  cur[r] = tok;
}

int greedy_update(const int64_t* next_ids, int64_t* cur_tokens, int64_t* out_tokens, int* out_len,
                  unsigned char* finished, int* n_finished, int R, int max_len, int pos, long long end_id,
                  long long pad_id, cudaStream_t s) {
  greedy_update_kernel<<<cdiv(R, 128), 128, 0, s>>>(next_ids, cur_tokens, out_tokens, out_len, finished, n_finished,
                                                    R, max_len, pos, end_id, pad_id);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// beam search step: one CTA per image.  Candidates: beam x V scores = parent score + log-softmax.
// A finished parent contributes a single candidate (END, same score).  Ties go to the lower flat
// index (parent-major, then token id).
// ------------------------------------------------------------------------------------------
static constexpr int TOPK_MAX = 4;

struct Cand { float s; int idx; };
__device__ __forceinline__ bool better(const Cand& a, const Cand& b) {
  return a.s > b.s || (a.s == b.s && a.idx < b.idx);
}
__device__ __forceinline__ void insert_topk(Cand (&top)[TOPK_MAX], Cand c, int k) {
#pragma unroll
  for (int i = 0; i < TOPK_MAX; ++i) {
    if (i < k && better(c, top[i])) {
      const Cand t = top[i];
      top[i] = c;
      c = t;
    }
  }
}

__global__ void __launch_bounds__(256)
beam_topk_kernel(const float* __restrict__ logits, const float* __restrict__ beam_scores,
                 const unsigned char* __restrict__ finished, int beam, int V, long long end_id, int first_step,
                 int64_t* __restrict__ out_tokens, int* __restrict__ out_parent, float* __restrict__ out_scores) {
  __shared__ float s_lse[TOPK_MAX];
  __shared__ float s_red[8];
  __shared__ Cand s_top[8][TOPK_MAX];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // log-sum-exp per parent row
  for (int p = 0; p < beam; ++p) {
    const float* row = logits + (static_cast<long long>(b) * beam + p) * V;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < V; j += blockDim.x) m = fmaxf(m, row[j]);
    m = warp_max(m);
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    m = s_red[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = threadIdx.x; j < V; j += blockDim.x) sum += expf(row[j] - m);
    sum = warp_sum(sum);
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += s_red[w];
      s_lse[p] = m + logf(t);
    }
    __syncthreads();
  }
  Cand top[TOPK_MAX];
#pragma unroll
  for (int i = 0; i < TOPK_MAX; ++i) { top[i].s = -INFINITY; top[i].idx = 0x7fffffff; }
  const int nparents = first_step ? 1 : beam;   // step 0: all beams are identical copies of START
  for (int p = 0; p < nparents; ++p) {
    const long long r = static_cast<long long>(b) * beam + p;
    const float base = beam_scores[r];
    if (finished[r]) {
      if (threadIdx.x == 0) insert_topk(top, Cand{base, p * V + static_cast<int>(end_id)}, beam);
      continue;
    }
    const float* row = logits + r * V;
    const float lse = s_lse[p];
    for (int j = threadIdx.x; j < V; j += blockDim.x) {
      insert_topk(top, Cand{base + (row[j] - lse), p * V + j}, beam);
    }
  }
  // warp merge, then CTA merge
  for (int off = 16; off > 0; off >>= 1) {
    Cand other[TOPK_MAX];
#pragma unroll
    for (int i = 0; i < TOPK_MAX; ++i) {
      other[i].s = __shfl_xor_sync(0xffffffffu, top[i].s, off);
      other[i].idx = __shfl_xor_sync(0xffffffffu, top[i].idx, off);
    }
#pragma unroll
    for (int i = 0; i < TOPK_MAX; ++i)
      if (i < beam) insert_topk(top, other[i], beam);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < TOPK_MAX; ++i) s_top[warp][i] = top[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      for (int i = 0; i < beam; ++i) insert_topk(top, s_top[w][i], beam);
    for (int i = 0; i < beam; ++i) {
      const long long r = static_cast<long long>(b) * beam + i;
      out_tokens[r] = top[i].idx % V;
      out_parent[r] = top[i].idx / V;
      out_scores[r] = top[i].s;
    }
  }
}

int beam_topk(const float* logits, const float* beam_scores, const unsigned char* finished, int B, int beam,
              int V, long long end_id, int first_step, int64_t* out_tokens, int* out_parent, float* out_scores,
              cudaStream_t s) {
  B200_REQUIRE(beam >= 1 && beam <= TOPK_MAX, "beam_topk: beam %d out of range (1..%d)", beam, TOPK_MAX);
  beam_topk_kernel<<<B, 256, 0, s>>>(logits, beam_scores, finished, beam, V, end_id, first_step, out_tokens,
                                     out_parent, out_scores);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// sequences / flags / scores follow their parents; the chosen token is appended at pos+1
__global__ void beam_advance_kernel(const int64_t* __restrict__ seq_in, int64_t* __restrict__ seq_out,
                                    const unsigned char* __restrict__ fin_in, unsigned char* __restrict__ fin_out,
                                    const int64_t* __restrict__ tokens, const int* __restrict__ parent, int R,
                                    int beam, int max_len, int pos, long long end_id) {
  const int r = blockIdx.x;
  if (r >= R) return;
  const int src = (r / beam) * beam + parent[r];
  for (int t = threadIdx.x; t <= pos; t += blockDim.x)
    seq_out[static_cast<long long>(r) * max_len + t] = seq_in[static_cast<long long>(src) * max_len + t];
  if (threadIdx.x == 0) {
    const long long tok = tokens[r];
    seq_out[static_cast<long long>(r) * max_len + pos + 1] = tok;
    fin_out[r] = (fin_in[src] || tok == end_id) ? 1 : 0;
  }
}

int beam_advance(const int64_t* seq_in, int64_t* seq_out, const unsigned char* fin_in, unsigned char* fin_out,
                 const int64_t* tokens, const int* parent, int R, int beam, int max_len, int pos, long long end_id,
                 cudaStream_t s) {
  beam_advance_kernel<<<R, 64, 0, s>>>(seq_in, seq_out, fin_in, fin_out, tokens, parent, R, beam, max_len, pos, end_id);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// best hypothesis per image (first index on score ties), cut after its first END, PAD afterwards
__global__ void beam_finalize_kernel(const int64_t* __restrict__ seqs, const float* __restrict__ scores, int beam,
                                     int max_len, int n_tok, long long end_id, long long pad_id,
                                     int64_t* __restrict__ out_tokens, int* __restrict__ out_len,
                                     float* __restrict__ out_score) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  int best = 0;
  for (int i = 1; i < beam; ++i)
    if (scores[b * beam + i] > scores[b * beam + best]) best = i;
  const int64_t* src = seqs + (static_cast<long long>(b) * beam + best) * max_len;
  int len = n_tok;
  for (int t = 1; t < n_tok; ++t)
    if (src[t] == end_id) { len = t + 1; break; }
  for (int t = 0; t < max_len; ++t) out_tokens[static_cast<long long>(b) * max_len + t] = t < len ? src[t] : pad_id;
  out_len[b] = len;
  if (out_score) out_score[b] = scores[b * beam + best];
}

int beam_finalize(const int64_t* seqs, const float* scores, int B, int beam, int max_len, int n_tok,
                  long long end_id, long long pad_id, int64_t* out_tokens, int* out_len, float* out_score,
                  cudaStream_t s) {
  beam_finalize_kernel<<<B, 32, 0, s>>>(seqs, scores, beam, max_len, n_tok, end_id, pad_id, out_tokens, out_len, out_score);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void fill_i64_kernel(int64_t* p, long long n, long long v) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}
int fill_i64(int64_t* p, long long n, long long v, cudaStream_t s) {
  fill_i64_kernel<<<cdiv(n, 256), 256, 0, s>>>(p, n, v);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// out[r*stride] = v for r < n   (column fill: START token of every sequence)
__global__ void fill_col_i64_kernel(int64_t* p, long long n, long long stride, long long v) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) p[i * stride] = v;
}
int fill_col_i64(int64_t* p, long long n, long long stride, long long v, cudaStream_t s) {
  fill_col_i64_kernel<<<cdiv(n, 256), 256, 0, s>>>(p, n, stride, v);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void cache_reorder_kernel(const bf16* __restrict__ src_k, const bf16* __restrict__ src_v,
                                     bf16* __restrict__ dst_k, bf16* __restrict__ dst_v,
                                     const int* __restrict__ parent, int beam, int H, int hd, int max_len, int npos) {
  const int vpr = hd >> 3;
  const int r = blockIdx.y;                      // destination row
  const int b = r / beam;
  const int src_r = b * beam + parent[r];
  const long long per_row = static_cast<long long>(H) * npos * vpr;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < per_row;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % vpr);
    const int t = static_cast<int>((idx / vpr) % npos);
    const int h = static_cast<int>(idx / (static_cast<long long>(vpr) * npos));
    const long long so = ((static_cast<long long>(src_r) * H + h) * max_len + t) * hd + c * 8;
    const long long dof = ((static_cast<long long>(r) * H + h) * max_len + t) * hd + c * 8;
    *reinterpret_cast<uint4*>(dst_k + dof) = *reinterpret_cast<const uint4*>(src_k + so);
    *reinterpret_cast<uint4*>(dst_v + dof) = *reinterpret_cast<const uint4*>(src_v + so);
  }
}

int cache_reorder(const bf16* src_k, const bf16* src_v, bf16* dst_k, bf16* dst_v, const int* parent, int B,
                  int beam, int H, int hd, int max_len, int pos, cudaStream_t s) {
  const int npos = pos + 1;
  const long long per_row = static_cast<long long>(H) * npos * (hd / 8);
  dim3 grid(cdiv(per_row, 256) > 8 ? 8 : cdiv(per_row, 256), B * beam);
  cache_reorder_kernel<<<grid, 256, 0, s>>>(src_k, src_v, dst_k, dst_v, parent, beam, H, hd, max_len, npos);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
