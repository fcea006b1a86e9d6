// Fused attention for the caption decoder (replaces F.scaled_dot_product_attention reached from
// torch/nn/functional.py:6682 via nn.MultiheadAttention):
//   self attention : causal (key j <= query i, utils.py:30-36) + key padding taken straight from
//                    the token ids (utils.py:66, decoder.py:158-162), never materialised as a mask;
//   cross attention: every query sees all S image tokens (model.py:158), optional key padding.
// One CTA per (image, head).  The whole K and V of that (image, head) — 50..257 ViT/CLIP tokens or
// <= 99 caption positions — are loaded ONCE into shared memory and stay resident while every
// query tile is processed; scores/probabilities live in registers (flash-style online softmax),
// products run on the warp-level tensor-core path (mma.sync m16n8k16 bf16, fp32 accumulate;
// the problems are 47x47..47x257 per head, far below one tcgen05 tile).
// Backward recomputes P from the saved log-sum-exp: phase 1 produces dQ per query tile, phase 2
// produces dK/dV per key tile (no atomics, deterministic).
#include "attention.cuh"
#include "attention_tc.cuh"
#include <math.h>
#include <stdlib.h>

namespace b200 {

static constexpr float LOG2E = 1.4426950408889634f;
static constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

struct AttnDev {
  const bf16 *q, *k, *v, *o_in, *d_o;
  bf16 *o, *dq, *dk, *dv;
  long long q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, o_bs, o_ts, do_bs, do_ts, dq_bs, dq_ts, dk_bs, dk_ts, dv_bs, dv_ts;
  float* lse;
  int B, H, Tq, Tk, TQP, TKP;
  int lse_ld, TqS, TkS;          // row pitch of lse and the (row, key) strides of the dropout counter: the LARGEST lengths
  const int32_t *cu_q, *cu_k;    // packed sequences (attention.cuh); the kernels then patch Tq / Tk / base pointers per sample
  int causal;
  const long long* key_tokens; long long pad_idx;
  const unsigned char* key_pad_mask;
  float scale;
  DropCfg drop;
};

// keep-decision of probability (row i, key j) of head bh; see common.cuh for the element order
__device__ __forceinline__ uint32_t attn_drop_pair(const AttnDev& p, int bh, int row, int key) {
  return (static_cast<uint32_t>(bh) * p.TqS + row) * static_cast<uint32_t>((p.TkS + 1) >> 1) + (static_cast<uint32_t>(key) >> 1);
}
// packed sequences: this CTA's sample owns rows [cu[b], cu[b+1]) -- Tq / Tk shrink to the sample's own lengths (the tile
// sizes TQP / TKP stay those of the longest sample).  The ROW OFFSETS are not patched into the struct: the kernels index
// the packed tensors as  base + bq * q_bs  with bq = cu_q[b] and q_bs = q_ts (set on the host), see VARLEN_ROWS.
// (A first version shifted the base pointers in the struct copy; nvcc 12.9 dropped the update of the FIRST field --
// later reads went back to the kernel parameter -- so the offsets are explicit locals now.)
__device__ __forceinline__ void attn_varlen_patch(AttnDev& p, int b) {
  if (p.cu_q) p.Tq = p.cu_q[b + 1] - p.cu_q[b];
  if (p.cu_k) p.Tk = p.cu_k[b + 1] - p.cu_k[b];
}
#define VARLEN_ROWS(p, b) const long long bq = (p).cu_q ? (p).cu_q[b] : (b), bk = (p).cu_k ? (p).cu_k[b] : (b); (void)bq; (void)bk
__device__ __forceinline__ bool attn_drop_keep(uint32_t r, int key, uint32_t thr) {
  return ((key & 1) ? (r >> 16) : (r & 0xFFFFu)) >= thr;
}

// cooperative asynchronous copy (cp.async, 16 B per request, all requests in flight at once) of
// `rows` x HD bf16 (row stride ts elements) into padded smem; rows [rows, rows_padded) are
// zero-filled through the src-size-0 form.  Completion: cp_async_wait_all() + __syncthreads().
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
template <int HD>
__device__ __forceinline__ void load_rows(bf16* dst, const bf16* src, long long ts, int rows, int rows_padded) {
  constexpr int LD = HD + 8;
  constexpr int VPR = HD / 8;
  for (int i = threadIdx.x; i < rows_padded * VPR; i += blockDim.x) {
    const int r = i / VPR, c = (i % VPR) * 8;
    const bool ok = r < rows;
    cp_async_16(dst + r * LD + c, src + (ok ? r * ts : 0) + c, ok ? 16 : 0);
  }
}

// C[16 x (NT*8)] += A[16 x HD] * B^T, A rows at a_row0 of sA, B rows (the "n" index) at b_row0 of
// sB, both stored [row][HD+8] (dimension contiguous).
template <int HD, int NT>
__device__ __forceinline__ void mma_abt(float (&c)[NT][4], const bf16* sA, int a_row0, const bf16* sB,
                                        int b_row0, int b_rows_avail) {
  constexpr int LD = HD + 8;
  const int lane = threadIdx.x & 31;
  const int mi = lane >> 3, r8 = lane & 7;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    uint32_t a[4];
    ldsm_x4(a, smem_u32(sA + (a_row0 + (mi & 1) * 8 + r8) * LD + ks * 16 + (mi >> 1) * 8));
#pragma unroll
    for (int j = 0; j < NT; j += 2) {
      if (j * 8 < b_rows_avail) {   // warp-uniform
        uint32_t b[4];
        ldsm_x4(b, smem_u32(sB + (b_row0 + (j + (mi >> 1)) * 8 + r8) * LD + ks * 16 + (mi & 1) * 8));
        mma16816(c[j], a, b[0], b[1]);
        mma16816(c[j + 1], a, b[2], b[3]);
      }
    }
  }
}

// C[16 x HD] += P[16 x (NT*8)] * B, P given as fp32 accumulator fragments (converted to bf16),
// B rows (the reduction index) at b_row0 of sB stored [row][HD+8].
template <int HD, int NT>
__device__ __forceinline__ void mma_pb(float (&c)[HD / 8][4], const float (&p)[NT][4], const bf16* sB,
                                       int b_row0, int b_rows_avail) {
  constexpr int LD = HD + 8;
  const int lane = threadIdx.x & 31;
  const int mi = lane >> 3, r8 = lane & 7;
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    if (kk * 16 < b_rows_avail) {   // warp-uniform
      uint32_t a[4];
      a[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
      a[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
      a[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
      a[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
      for (int dn = 0; dn < HD / 8; dn += 2) {
        uint32_t b[4];
        ldsm_x4_t(b, smem_u32(sB + (b_row0 + kk * 16 + (mi & 1) * 8 + r8) * LD + (dn + (mi >> 1)) * 8));
        mma16816(c[dn], a, b[0], b[1]);
        mma16816(c[dn + 1], a, b[2], b[3]);
      }
    }
  }
}

template <int HD>
struct AttnSmem {
  static constexpr int LD = HD + 8;
  static size_t fwd_bytes(int TQP, int TKP) {
    return static_cast<size_t>(TQP + 2 * TKP) * LD * 2 + TKP * sizeof(float) + 16;
  }
  static size_t bwd_bytes(int TQP, int TKP) {
    return static_cast<size_t>(3 * TQP + 2 * TKP) * LD * 2 + (TKP + 2 * TQP) * sizeof(float) + 16;
  }
};

__device__ __forceinline__ void fill_key_bias(float* sBias, const AttnDev& p, int b) {
  for (int j = threadIdx.x; j < p.TKP; j += blockDim.x) {
    bool masked = j >= p.Tk;
    if (!masked && p.key_tokens) masked = p.key_tokens[static_cast<long long>(b) * p.Tk + j] == p.pad_idx;
    if (!masked && p.key_pad_mask) masked = p.key_pad_mask[static_cast<long long>(b) * p.Tk + j] != 0;
    sBias[j] = masked ? -INFINITY : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int HD, int NT>
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const AttnDev pin) {
  pdl_wait();
  pdl_trigger();
  AttnDev p = pin;
  attn_varlen_patch(p, blockIdx.y);
  constexpr int LD = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
  bf16* sK = sQ + p.TQP * LD;
  bf16* sV = sK + p.TKP * LD;
  float* sBias = reinterpret_cast<float*>(sV + p.TKP * LD);

  const int h = blockIdx.x, b = blockIdx.y;
  VARLEN_ROWS(p, b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  load_rows<HD>(sQ, p.q + bq * p.q_bs + h * HD, p.q_ts, p.Tq, p.TQP);
  load_rows<HD>(sK, p.k + bk * p.k_bs + h * HD, p.k_ts, p.Tk, p.TKP);
  load_rows<HD>(sV, p.v + bk * p.v_bs + h * HD, p.v_ts, p.Tk, p.TKP);
  fill_key_bias(sBias, p, b);
  cp_async_wait_all();
  __syncthreads();

  const float sl2 = p.scale * LOG2E;
  const uint32_t dkey = p.drop.thr ? drop_key(p.drop) : 0u;
  const int bh = b * p.H + h;
  for (int mt = warp; mt * 16 < p.TQP; mt += 4) {
    const int row0 = mt * 16;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    const int kend = p.causal ? min(p.TKP, row0 + 16) : p.TKP;
    for (int kb0 = 0; kb0 < kend; kb0 += NT * 8) {
      float s[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
      const int avail = kend - kb0;
      mma_abt<HD, NT>(s, sQ, row0, sK, kb0, avail);
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kb0 + j * 8 + 2 * t + (e & 1);
          const int row = row0 + g + (e >> 1) * 8;
          float val = -INFINITY;
          if (j * 8 < avail && key < p.TKP && !(p.causal && key > row)) val = s[j][e] * sl2 + sBias[key];
          s[j][e] = val;
          mx[e >> 1] = fmaxf(mx[e >> 1], val);
        }
      }
      float alpha[2], m_new[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        m_new[r] = fmaxf(m_run[r], quad_max(mx[r]));
        alpha[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f(m_run[r] - m_new[r]);
        m_run[r] = m_new[r];
      }
      float ps[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float mm = m_new[e >> 1];
          const float pv = (mm == -INFINITY) ? 0.f : exp2f(s[j][e] - mm);
          s[j][e] = pv;
          ps[e >> 1] += pv;
        }
        if (p.drop.thr) {      // the row sum keeps the undropped probabilities; P V sees dropout(P)
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
            const uint32_t r = drop_rand(dkey, attn_drop_pair(p, bh, row0 + g + rh * 8, kb0 + j * 8 + 2 * t));
            drop_apply2(s[j][2 * rh], s[j][2 * rh + 1], r, p.drop.thr, p.drop.scale);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + ps[r];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        o[i][0] *= alpha[0]; o[i][1] *= alpha[0]; o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
      }
      mma_pb<HD, NT>(o, s, sV, kb0, avail);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float l = quad_sum(l_run[r]);
      const int row = row0 + g + r * 8;
      const float inv = l > 0.f ? 1.f / l : 0.f;
      if (row < p.Tq) {
        bf16* orow = p.o + bq * p.o_bs + row * p.o_ts + h * HD;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i)
          *reinterpret_cast<uint32_t*>(orow + i * 8 + 2 * t) = pack_bf16(o[i][2 * r] * inv, o[i][2 * r + 1] * inv);
        if (p.lse && t == 0)
          p.lse[(static_cast<long long>(b) * p.H + h) * p.lse_ld + row] = (l > 0.f) ? m_run[r] * LN2 + logf(l) : -INFINITY;
      }
    }
  }
}

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------
// forward, split-key variant (the default when it fits): the keys of one (image, head) are cut into
// KS ranges of whole 16-key tiles and every (query tile, key range) pair gets its OWN warp, so a
// 47 x 197 problem keeps 12 warps busy instead of 3.  Each warp does one S = Q K^T block (<= NT*8
// keys), a local softmax and one P V product; the KS partial (max, sum, O) triples of a query tile are
// merged through shared memory (aliased onto the K/V tiles once every warp is done with them).
// The elementwise part is written for issue slots: key bias (padding) in registers, causal
// compare only where the block touches the diagonal, ex2.approx.ftz, no per-element branches.
// ------------------------------------------------------------------------------------------
template <int HD, int NT>
__global__ void __launch_bounds__(512)
attn_fwd_split_kernel(const AttnDev pin, int KS) {
  pdl_wait();
  pdl_trigger();
  AttnDev p = pin;
  attn_varlen_patch(p, blockIdx.y);
  constexpr int LD = HD + 8;
  constexpr int LDO = HD + 4;                       // fp32 partial-O row pitch
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
  bf16* sK = sQ + p.TQP * LD;
  bf16* sV = sK + p.TKP * LD;
  float* sBias = reinterpret_cast<float*>(sV + p.TKP * LD);
  float* sML = sBias + p.TKP;                       // [items][16][2] local (max, sum)
  float* sO = reinterpret_cast<float*>(sK);         // [items][16][LDO] partial O, aliases K/V

  const int h = blockIdx.x, b = blockIdx.y;
  VARLEN_ROWS(p, b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n_kt = p.TKP / 16;

  load_rows<HD>(sQ, p.q + bq * p.q_bs + h * HD, p.q_ts, p.Tq, p.TQP);
  load_rows<HD>(sK, p.k + bk * p.k_bs + h * HD, p.k_ts, p.Tk, p.TKP);
  load_rows<HD>(sV, p.v + bk * p.v_bs + h * HD, p.v_ts, p.Tk, p.TKP);
  fill_key_bias(sBias, p, b);
  cp_async_wait_all();
  __syncthreads();

  const int mt = warp / KS, ks = warp % KS;
  const int row0 = mt * 16;
  int kt0 = (n_kt * ks) / KS, kt1 = (n_kt * (ks + 1)) / KS;
  if (p.causal) kt1 = min(kt1, mt + 1);             // key tiles right of the diagonal are fully masked
  const int key0 = kt0 * 16;
  const int avail = max(kt1 - kt0, 0) * 16;         // keys of this item (multiple of 16, <= NT * 8)

  float s[NT][4], o[HD / 8][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_loc[2] = {-INFINITY, -INFINITY}, l_loc[2] = {0.f, 0.f};
  if (avail > 0) {
    mma_abt<HD, NT>(s, sQ, row0, sK, key0, avail);
    const float sl2 = p.scale * LOG2E;
    const bool diag = p.causal && (key0 + avail > row0);     // warp-uniform: block reaches the diagonal
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      if (j * 8 < avail) {                                    // warp-uniform
        const float2 kb = *reinterpret_cast<const float2*>(sBias + key0 + j * 8 + 2 * t);
        s[j][0] = fmaf(s[j][0], sl2, kb.x); s[j][1] = fmaf(s[j][1], sl2, kb.y);
        s[j][2] = fmaf(s[j][2], sl2, kb.x); s[j][3] = fmaf(s[j][3], sl2, kb.y);
        if (diag) {
          const int key = key0 + j * 8 + 2 * t, r0 = row0 + g;
          if (key > r0) s[j][0] = -INFINITY;
          if (key + 1 > r0) s[j][1] = -INFINITY;
          if (key > r0 + 8) s[j][2] = -INFINITY;
          if (key + 1 > r0 + 8) s[j][3] = -INFINITY;
        }
        mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
      } else {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = -INFINITY;
      }
    }
    m_loc[0] = quad_max(mx[0]);
    m_loc[1] = quad_max(mx[1]);
    const float ms0 = (m_loc[0] == -INFINITY) ? 0.f : m_loc[0];   // fully masked row: every p becomes 0
    const float ms1 = (m_loc[1] == -INFINITY) ? 0.f : m_loc[1];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] = ex2_ftz(s[j][0] - ms0); s[j][1] = ex2_ftz(s[j][1] - ms0);
      s[j][2] = ex2_ftz(s[j][2] - ms1); s[j][3] = ex2_ftz(s[j][3] - ms1);
      l_loc[0] += s[j][0] + s[j][1];
      l_loc[1] += s[j][2] + s[j][3];
    }
    l_loc[0] = quad_sum(l_loc[0]);
    l_loc[1] = quad_sum(l_loc[1]);
    if (p.drop.thr) {      // the row sums keep the undropped probabilities; P V sees dropout(P)
      const uint32_t dkey = drop_key(p.drop);
      const int bh = b * p.H + h;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if (j * 8 < avail) {
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
            const uint32_t r = drop_rand(dkey, attn_drop_pair(p, bh, row0 + g + rh * 8, key0 + j * 8 + 2 * t));
            drop_apply2(s[j][2 * rh], s[j][2 * rh + 1], r, p.drop.thr, p.drop.scale);
          }
        }
      }
    }
    mma_pb<HD, NT>(o, s, sV, key0, avail);
  }

  if (KS == 1) {           // no merge needed: normalise and store straight from registers
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = row0 + g + r * 8;
      const float inv = l_loc[r] > 0.f ? 1.f / l_loc[r] : 0.f;
      if (row < p.Tq) {
        bf16* orow = p.o + bq * p.o_bs + row * p.o_ts + h * HD;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i)
          *reinterpret_cast<uint32_t*>(orow + i * 8 + 2 * t) = pack_bf16(o[i][2 * r] * inv, o[i][2 * r + 1] * inv);
        if (p.lse && t == 0)
          p.lse[(static_cast<long long>(b) * p.H + h) * p.lse_ld + row] = (l_loc[r] > 0.f) ? m_loc[r] * LN2 + logf(l_loc[r]) : -INFINITY;
      }
    }
    return;
  }

  __syncthreads();         // every warp is done with K and V: their space now holds the partial O tiles
  {
    float* po = sO + static_cast<size_t>(warp) * 16 * LDO;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      *reinterpret_cast<float2*>(po + g * LDO + i * 8 + 2 * t) = make_float2(o[i][0], o[i][1]);
      *reinterpret_cast<float2*>(po + (g + 8) * LDO + i * 8 + 2 * t) = make_float2(o[i][2], o[i][3]);
    }
    if (t == 0) {
      sML[(warp * 16 + g) * 2] = m_loc[0]; sML[(warp * 16 + g) * 2 + 1] = l_loc[0];
      sML[(warp * 16 + g + 8) * 2] = m_loc[1]; sML[(warp * 16 + g + 8) * 2 + 1] = l_loc[1];
    }
  }
  __syncthreads();
  // merge: the KS warps of a query tile share its 16 rows; each lane owns one 8-wide column strip of a row
  constexpr int CPR = HD / 8;                      // 8-column strips per row
  const int first = mt * KS;                       // first item of this query tile
  for (int idx = ks * 32 + lane; idx < 16 * CPR; idx += KS * 32) {
    const int rr = idx / CPR, cs = (idx % CPR) * 8;
    float M = -INFINITY;
    for (int q = 0; q < KS; ++q) M = fmaxf(M, sML[((first + q) * 16 + rr) * 2]);
    float L = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < KS; ++q) {
      const float mq = sML[((first + q) * 16 + rr) * 2];
      const float w = (mq == -INFINITY) ? 0.f : ex2_ftz(mq - M);
      L = fmaf(sML[((first + q) * 16 + rr) * 2 + 1], w, L);
      const float* src = sO + (static_cast<size_t>(first + q) * 16 + rr) * LDO + cs;
      const float4 a0 = *reinterpret_cast<const float4*>(src), a1 = *reinterpret_cast<const float4*>(src + 4);
      acc[0] = fmaf(a0.x, w, acc[0]); acc[1] = fmaf(a0.y, w, acc[1]); acc[2] = fmaf(a0.z, w, acc[2]); acc[3] = fmaf(a0.w, w, acc[3]);
      acc[4] = fmaf(a1.x, w, acc[4]); acc[5] = fmaf(a1.y, w, acc[5]); acc[6] = fmaf(a1.z, w, acc[6]); acc[7] = fmaf(a1.w, w, acc[7]);
    }
    const int row = row0 + rr;
    if (row < p.Tq) {
      const float inv = L > 0.f ? 1.f / L : 0.f;
      uint4 out;
      out.x = pack_bf16(acc[0] * inv, acc[1] * inv); out.y = pack_bf16(acc[2] * inv, acc[3] * inv);
      out.z = pack_bf16(acc[4] * inv, acc[5] * inv); out.w = pack_bf16(acc[6] * inv, acc[7] * inv);
      *reinterpret_cast<uint4*>(p.o + bq * p.o_bs + row * p.o_ts + h * HD + cs) = out;
      if (p.lse && cs == 0)
        p.lse[(static_cast<long long>(b) * p.H + h) * p.lse_ld + row] = (L > 0.f) ? M * LN2 + logf(L) : -INFINITY;
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
template <int HD, int NT>
__global__ void __launch_bounds__(128)
attn_bwd_kernel(const AttnDev pin) {
  pdl_wait();
  pdl_trigger();
  AttnDev p = pin;
  attn_varlen_patch(p, blockIdx.y);
  constexpr int LD = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
  bf16* sdO = sQ + p.TQP * LD;
  bf16* sK = sdO + p.TQP * LD;
  bf16* sV = sK + p.TKP * LD;
  bf16* sO = sV + p.TKP * LD;
  float* sBias = reinterpret_cast<float*>(sO + p.TQP * LD);
  float* sLse = sBias + p.TKP;   // already multiplied by log2(e); +inf on padded query rows
  float* sD = sLse + p.TQP;

  const int h = blockIdx.x, b = blockIdx.y;
  VARLEN_ROWS(p, b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  load_rows<HD>(sQ, p.q + bq * p.q_bs + h * HD, p.q_ts, p.Tq, p.TQP);
  load_rows<HD>(sdO, p.d_o + bq * p.do_bs + h * HD, p.do_ts, p.Tq, p.TQP);
  load_rows<HD>(sK, p.k + bk * p.k_bs + h * HD, p.k_ts, p.Tk, p.TKP);
  load_rows<HD>(sV, p.v + bk * p.v_bs + h * HD, p.v_ts, p.Tk, p.TKP);
  load_rows<HD>(sO, p.o_in + bq * p.o_bs + h * HD, p.o_ts, p.Tq, p.TQP);
  fill_key_bias(sBias, p, b);
  for (int row = threadIdx.x; row < p.TQP; row += blockDim.x) {
    float l = INFINITY;
    if (row < p.Tq) {
      l = p.lse[(static_cast<long long>(b) * p.H + h) * p.lse_ld + row];
      l = (l == -INFINITY) ? INFINITY : l * LOG2E;   // fully masked row: P = 0
    }
    sLse[row] = l;
  }
  cp_async_wait_all();
  __syncthreads();
  // D_i = sum_d dO[i,d] * O[i,d]; one warp per row, from shared memory
  for (int row = warp; row < p.TQP; row += 4) {
    float acc = 0.f;
    for (int c = lane * 2; c < HD; c += 64) {
      const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(sO + row * LD + c));
      const float2 d = unpack_bf16(*reinterpret_cast<const uint32_t*>(sdO + row * LD + c));
      acc += a.x * d.x + a.y * d.y;
    }
    acc = warp_sum(acc);
    if (lane == 0) sD[row] = acc;
  }
  __syncthreads();

  const float sl2 = p.scale * LOG2E;
  const uint32_t dkey = p.drop.thr ? drop_key(p.drop) : 0u;
  const int bh = b * p.H + h;

  // ---- phase 1: dQ, one 16-query tile per warp iteration
  for (int mt = warp; mt * 16 < p.TQP; mt += 4) {
    const int row0 = mt * 16;
    float dq[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
    const float lse0 = sLse[row0 + g], lse1 = sLse[row0 + g + 8];
    const float d0 = sD[row0 + g], d1 = sD[row0 + g + 8];
    const int kend = p.causal ? min(p.TKP, row0 + 16) : p.TKP;
    for (int kb0 = 0; kb0 < kend; kb0 += NT * 8) {
      float s[NT][4], dp[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
      }
      const int avail = kend - kb0;
      mma_abt<HD, NT>(s, sQ, row0, sK, kb0, avail);
      mma_abt<HD, NT>(dp, sdO, row0, sV, kb0, avail);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kb0 + j * 8 + 2 * t + (e & 1);
          const int row = row0 + g + (e >> 1) * 8;
          float ds = 0.f;
          if (j * 8 < avail && key < p.TKP && !(p.causal && key > row)) {
            const float pv = exp2f(s[j][e] * sl2 + sBias[key] - ((e >> 1) ? lse1 : lse0));
            float dpe = dp[j][e];
            if (p.drop.thr)
              dpe = attn_drop_keep(drop_rand(dkey, attn_drop_pair(p, bh, row, key)), key, p.drop.thr) ? dpe * p.drop.scale : 0.f;
            ds = pv * (dpe - ((e >> 1) ? d1 : d0)) * p.scale;
          }
          s[j][e] = ds;
        }
      }
      mma_pb<HD, NT>(dq, s, sK, kb0, avail);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = row0 + g + r * 8;
      if (row < p.Tq) {
        bf16* drow = p.dq + bq * p.dq_bs + row * p.dq_ts + h * HD;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i)
          *reinterpret_cast<uint32_t*>(drow + i * 8 + 2 * t) = pack_bf16(dq[i][2 * r], dq[i][2 * r + 1]);
      }
    }
  }

  // ---- phase 2: dK, dV, one 16-key tile per warp iteration (scores computed transposed)
  for (int kt = warp; kt * 16 < p.TKP; kt += 4) {
    const int key0 = kt * 16;
    float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    const float kb_0 = sBias[key0 + g], kb_1 = sBias[key0 + g + 8];
    // causal: only queries >= key contribute; start at the 16-aligned query tile holding key0
    const int qstart = p.causal ? min(key0, p.TQP) : 0;
    for (int q0 = qstart; q0 < p.TQP; q0 += NT * 8) {
      float st[NT][4], dpt[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
        dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
      }
      const int avail = p.TQP - q0;
      mma_abt<HD, NT>(st, sK, key0, sQ, q0, avail);
      mma_abt<HD, NT>(dpt, sV, key0, sdO, q0, avail);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qi = q0 + j * 8 + 2 * t + (e & 1);
          const int key = key0 + g + (e >> 1) * 8;
          float pv = 0.f, ds = 0.f;
          if (j * 8 < avail && qi < p.TQP && !(p.causal && key > qi)) {
            pv = exp2f(st[j][e] * sl2 + ((e >> 1) ? kb_1 : kb_0) - sLse[qi]);
            float dpe = dpt[j][e];
            bool keep = true;
            if (p.drop.thr) {
              keep = attn_drop_keep(drop_rand(dkey, attn_drop_pair(p, bh, qi, key)), key, p.drop.thr);
              dpe = keep ? dpe * p.drop.scale : 0.f;
            }
            ds = pv * (dpe - sD[qi]) * p.scale;
            if (p.drop.thr) pv = keep ? pv * p.drop.scale : 0.f;     // dV = dropout(P)^T dO
          }
          st[j][e] = pv;
          dpt[j][e] = ds;
        }
      }
      mma_pb<HD, NT>(dv, st, sdO, q0, avail);
      mma_pb<HD, NT>(dk, dpt, sQ, q0, avail);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int key = key0 + g + r * 8;
      if (key < p.Tk) {
        bf16* krow = p.dk + bk * p.dk_bs + key * p.dk_ts + h * HD;
        bf16* vrow = p.dv + bk * p.dv_bs + key * p.dv_ts + h * HD;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
          *reinterpret_cast<uint32_t*>(krow + i * 8 + 2 * t) = pack_bf16(dk[i][2 * r], dk[i][2 * r + 1]);
          *reinterpret_cast<uint32_t*>(vrow + i * 8 + 2 * t) = pack_bf16(dv[i][2 * r], dv[i][2 * r + 1]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward, staged variant (the default): S and dP are computed ONCE per (query tile, key block)
// work item, P and dS go through shared memory in bf16 so that every later product can fetch the
// operand orientation it needs with ldmatrix(.trans):
//   step 1   items (q-tile, key block): S = Q K^T, dP = dO V^T -> P, dS kept packed in registers
//   step 2a  P -> smem;  dV[key tile] = P^T dO
//   step 2b  dS -> same smem;  dK[key tile] = dS^T Q  and  dQ[q-tile, half of hd] = dS K
// 5 products instead of 7, 8 balanced warps, <= 128 registers, two CTAs per SM.
// ------------------------------------------------------------------------------------------
template <int HD, int NT, int MAXI, int MINB, int NW>
__global__ void __launch_bounds__(NW * 32, MINB)
attn_bwd2_kernel(const AttnDev pin) {
  pdl_wait();
  pdl_trigger();
  AttnDev p = pin;
  attn_varlen_patch(p, blockIdx.y);
  constexpr int LD = HD + 8;
  constexpr int KB = NT * 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int LDP = p.TKP + 8;
  bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
  bf16* sdO = sQ + p.TQP * LD;
  bf16* sK = sdO + p.TQP * LD;
  bf16* sV = sK + p.TKP * LD;
  bf16* sPS = sV + p.TKP * LD;                       // [TQP][LDP]; first holds O (for D), then P, then dS
  const int ps_elems = p.TQP * (LDP > LD ? LDP : LD);
  float* sBias = reinterpret_cast<float*>(sPS + ps_elems);
  float* sLse = sBias + p.TKP;
  float* sD = sLse + p.TQP;

  const int h = blockIdx.x, b = blockIdx.y;
  VARLEN_ROWS(p, b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;

  load_rows<HD>(sQ, p.q + bq * p.q_bs + h * HD, p.q_ts, p.Tq, p.TQP);
  load_rows<HD>(sdO, p.d_o + bq * p.do_bs + h * HD, p.do_ts, p.Tq, p.TQP);
  load_rows<HD>(sK, p.k + bk * p.k_bs + h * HD, p.k_ts, p.Tk, p.TKP);
  load_rows<HD>(sV, p.v + bk * p.v_bs + h * HD, p.v_ts, p.Tk, p.TKP);
  load_rows<HD>(sPS, p.o_in + bq * p.o_bs + h * HD, p.o_ts, p.Tq, p.TQP);   // O, row pitch LD
  fill_key_bias(sBias, p, b);
  for (int row = threadIdx.x; row < p.TQP; row += blockDim.x) {
    float l = INFINITY;
    if (row < p.Tq) {
      l = p.lse[(static_cast<long long>(b) * p.H + h) * p.lse_ld + row];
      l = (l == -INFINITY) ? INFINITY : l * LOG2E;
    }
    sLse[row] = l;
  }
  cp_async_wait_all();
  __syncthreads();
  for (int row = warp; row < p.TQP; row += NW) {
    float acc = 0.f;
    for (int c = lane * 2; c < HD; c += 64) {
      const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(sPS + row * LD + c));
      const float2 d = unpack_bf16(*reinterpret_cast<const uint32_t*>(sdO + row * LD + c));
      acc += a.x * d.x + a.y * d.y;
    }
    acc = warp_sum(acc);
    if (lane == 0) sD[row] = acc;
  }
  __syncthreads();   // O is dead from here on; sPS is free

  const float sl2 = p.scale * LOG2E;
  const uint32_t dkey = p.drop.thr ? drop_key(p.drop) : 0u;
  const int bh = b * p.H + h;
  const int n_qt = p.TQP / 16;
  const int NB = (p.TKP + KB - 1) / KB;
  const int items = n_qt * NB;

  // ---- step 1
  uint32_t pP[MAXI][NT][2], pS[MAXI][NT][2];
#pragma unroll
  for (int ii = 0; ii < MAXI; ++ii) {
#pragma unroll
    for (int j = 0; j < NT; ++j) { pP[ii][j][0] = pP[ii][j][1] = 0u; pS[ii][j][0] = pS[ii][j][1] = 0u; }
    const int item = warp + ii * NW;
    if (item < items) {
      const int mt = item / NB, kb = item % NB;
      const int row0 = mt * 16, key0 = kb * KB;
      if (!(p.causal && key0 > row0 + 15)) {
        const int avail = min(KB, p.TKP - key0);
        float sacc[NT][4], dp[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.f;
          dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
        }
        mma_abt<HD, NT>(sacc, sQ, row0, sK, key0, avail);
        mma_abt<HD, NT>(dp, sdO, row0, sV, key0, avail);
        const float lse0 = sLse[row0 + g], lse1 = sLse[row0 + g + 8];
        const float d0 = sD[row0 + g], d1 = sD[row0 + g + 8];
        const bool diag = p.causal && (key0 + avail > row0);      // warp-uniform: block reaches the diagonal
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          float pv[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f};
          if (j * 8 < avail) {                                   // warp-uniform
            // masked keys carry a -inf bias, padded query rows a +inf lse: both give P = 0 without a branch
            const float2 kb = *reinterpret_cast<const float2*>(sBias + key0 + j * 8 + 2 * t);
            pv[0] = ex2_ftz(fmaf(sacc[j][0], sl2, kb.x) - lse0);
            pv[1] = ex2_ftz(fmaf(sacc[j][1], sl2, kb.y) - lse0);
            pv[2] = ex2_ftz(fmaf(sacc[j][2], sl2, kb.x) - lse1);
            pv[3] = ex2_ftz(fmaf(sacc[j][3], sl2, kb.y) - lse1);
            if (diag) {
              const int key = key0 + j * 8 + 2 * t, r0 = row0 + g;
              if (key > r0) pv[0] = 0.f;
              if (key + 1 > r0) pv[1] = 0.f;
              if (key > r0 + 8) pv[2] = 0.f;
              if (key + 1 > r0 + 8) pv[3] = 0.f;
            }
            float dpe[4] = {dp[j][0], dp[j][1], dp[j][2], dp[j][3]};
            if (p.drop.thr) {
              const uint32_t ra = drop_rand(dkey, attn_drop_pair(p, bh, row0 + g, key0 + j * 8 + 2 * t));
              const uint32_t rb = drop_rand(dkey, attn_drop_pair(p, bh, row0 + g + 8, key0 + j * 8 + 2 * t));
              drop_apply2(dpe[0], dpe[1], ra, p.drop.thr, p.drop.scale);
              drop_apply2(dpe[2], dpe[3], rb, p.drop.thr, p.drop.scale);
              ds[0] = pv[0] * (dpe[0] - d0) * p.scale; ds[1] = pv[1] * (dpe[1] - d0) * p.scale;
              ds[2] = pv[2] * (dpe[2] - d1) * p.scale; ds[3] = pv[3] * (dpe[3] - d1) * p.scale;
              drop_apply2(pv[0], pv[1], ra, p.drop.thr, p.drop.scale);      // dV = dropout(P)^T dO
              drop_apply2(pv[2], pv[3], rb, p.drop.thr, p.drop.scale);
            } else {
              ds[0] = pv[0] * (dpe[0] - d0) * p.scale; ds[1] = pv[1] * (dpe[1] - d0) * p.scale;
              ds[2] = pv[2] * (dpe[2] - d1) * p.scale; ds[3] = pv[3] * (dpe[3] - d1) * p.scale;
            }
          }
          pP[ii][j][0] = pack_bf16(pv[0], pv[1]); pP[ii][j][1] = pack_bf16(pv[2], pv[3]);
          pS[ii][j][0] = pack_bf16(ds[0], ds[1]); pS[ii][j][1] = pack_bf16(ds[2], ds[3]);
        }
      }
    }
  }

  auto stage = [&](const uint32_t (&src)[MAXI][NT][2]) {
#pragma unroll
    for (int ii = 0; ii < MAXI; ++ii) {
      const int item = warp + ii * NW;
      if (item < items) {
        const int mt = item / NB, kb = item % NB;
        const int row0 = mt * 16, key0 = kb * KB;
        const int avail = min(KB, p.TKP - key0);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          if (j * 8 < avail) {
            *reinterpret_cast<uint32_t*>(sPS + (row0 + g) * LDP + key0 + j * 8 + 2 * t) = src[ii][j][0];
            *reinterpret_cast<uint32_t*>(sPS + (row0 + g + 8) * LDP + key0 + j * 8 + 2 * t) = src[ii][j][1];
          }
        }
      }
    }
  };

  // ---- step 2a: dV[key tile] = P^T dO
  stage(pP);
  __syncthreads();
  const int n_kt = p.TKP / 16;
  for (int kt = warp; kt < n_kt; kt += NW) {
    float acc[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    const int qs0 = p.causal ? min(kt, n_qt) : 0;
    for (int qs = qs0; qs < n_qt; ++qs) {
      uint32_t a[4];
      ldsm_x4_t(a, smem_u32(sPS + (qs * 16 + (mi >> 1) * 8 + r8) * LDP + kt * 16 + (mi & 1) * 8));
#pragma unroll
      for (int dn = 0; dn < HD / 8; dn += 2) {
        uint32_t bb[4];
        ldsm_x4_t(bb, smem_u32(sdO + (qs * 16 + (mi & 1) * 8 + r8) * LD + (dn + (mi >> 1)) * 8));
        mma16816(acc[dn], a, bb[0], bb[1]);
        mma16816(acc[dn + 1], a, bb[2], bb[3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int key = kt * 16 + g + r * 8;
      if (key < p.Tk) {
        bf16* vrow = p.dv + bk * p.dv_bs + key * p.dv_ts + h * HD;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i)
          *reinterpret_cast<uint32_t*>(vrow + i * 8 + 2 * t) = pack_bf16(acc[i][2 * r], acc[i][2 * r + 1]);
      }
    }
  }
  __syncthreads();

  // ---- step 2b: dS -> smem; dK[key tile] = dS^T Q; dQ[q tile, hd half] = dS K
  stage(pS);
  __syncthreads();
  const int n_dq = n_qt * 2;
  for (int w = warp; w < n_dq + n_kt; w += NW) {
    if (w < n_dq) {
      const int mt = w >> 1, half = w & 1;
      constexpr int NH = HD / 16;                  // n-tiles in one half of the head dim
      float acc[NH][4];
#pragma unroll
      for (int i = 0; i < NH; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
      const int ks_end = p.causal ? min(n_kt, mt + 1) : n_kt;
      for (int ks = 0; ks < ks_end; ++ks) {
        uint32_t a[4];
        ldsm_x4(a, smem_u32(sPS + (mt * 16 + (mi & 1) * 8 + r8) * LDP + ks * 16 + (mi >> 1) * 8));
#pragma unroll
        for (int dn = 0; dn < NH; dn += 2) {
          uint32_t bb[4];
          ldsm_x4_t(bb, smem_u32(sK + (ks * 16 + (mi & 1) * 8 + r8) * LD + (half * NH + dn + (mi >> 1)) * 8));
          mma16816(acc[dn], a, bb[0], bb[1]);
          mma16816(acc[dn + 1], a, bb[2], bb[3]);
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = mt * 16 + g + r * 8;
        if (row < p.Tq) {
          bf16* drow = p.dq + bq * p.dq_bs + row * p.dq_ts + h * HD + half * (HD / 2);
#pragma unroll
          for (int i = 0; i < NH; ++i)
            *reinterpret_cast<uint32_t*>(drow + i * 8 + 2 * t) = pack_bf16(acc[i][2 * r], acc[i][2 * r + 1]);
        }
      }
    } else {
      const int kt = w - n_dq;
      float acc[HD / 8][4];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
      const int qs0 = p.causal ? min(kt, n_qt) : 0;
      for (int qs = qs0; qs < n_qt; ++qs) {
        uint32_t a[4];
        ldsm_x4_t(a, smem_u32(sPS + (qs * 16 + (mi >> 1) * 8 + r8) * LDP + kt * 16 + (mi & 1) * 8));
#pragma unroll
        for (int dn = 0; dn < HD / 8; dn += 2) {
          uint32_t bb[4];
          ldsm_x4_t(bb, smem_u32(sQ + (qs * 16 + (mi & 1) * 8 + r8) * LD + (dn + (mi >> 1)) * 8));
          mma16816(acc[dn], a, bb[0], bb[1]);
          mma16816(acc[dn + 1], a, bb[2], bb[3]);
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int key = kt * 16 + g + r * 8;
        if (key < p.Tk) {
          bf16* krow = p.dk + bk * p.dk_bs + key * p.dk_ts + h * HD;
#pragma unroll
          for (int i = 0; i < HD / 8; ++i)
            *reinterpret_cast<uint32_t*>(krow + i * 8 + 2 * t) = pack_bf16(acc[i][2 * r], acc[i][2 * r + 1]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
static int check_common(const AttnArgs& a) {
  B200_REQUIRE(a.B > 0 && a.H > 0 && a.Tq > 0 && a.Tk > 0, "attention: empty problem");
  B200_REQUIRE(a.hd == 32 || a.hd == 64 || a.hd == 96 || a.hd == 128, "attention: head dim %d not in {32,64,96,128}", a.hd);
  B200_REQUIRE(a.Tq <= 512 && a.Tk <= 512, "attention: Tq/Tk (%d/%d) above the 512 resident-tile limit", a.Tq, a.Tk);
  B200_REQUIRE(!(a.cu_k && (a.key_tokens || a.key_pad_mask)), "attention: packed keys carry no padding, a key mask cannot be combined with cu_k");
  auto ok = [](const void* ptr, long long bs, long long ts) {
    return ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && bs % 8 == 0 && ts % 8 == 0;
  };
  B200_REQUIRE(ok(a.q, a.q_bs, a.q_ts) && ok(a.k, a.k_bs, a.k_ts) && ok(a.v, a.v_bs, a.v_ts) && ok(a.o, a.o_bs, a.o_ts),
               "attention: q/k/v/o must be non-null, 16-byte aligned, strides multiples of 8");
  return 0;
}

static void fill_dev(const AttnArgs& a, AttnDev* d) {
  memset(d, 0, sizeof(*d));
  d->q = a.q; d->k = a.k; d->v = a.v; d->o = a.o; d->o_in = a.o;
  d->q_bs = a.q_bs; d->q_ts = a.q_ts; d->k_bs = a.k_bs; d->k_ts = a.k_ts;
  d->v_bs = a.v_bs; d->v_ts = a.v_ts; d->o_bs = a.o_bs; d->o_ts = a.o_ts;
  d->lse = a.lse;
  d->B = a.B; d->H = a.H; d->Tq = a.Tq; d->Tk = a.Tk;
  d->TQP = (a.Tq + 15) / 16 * 16; d->TKP = (a.Tk + 15) / 16 * 16;
  d->lse_ld = a.Tq; d->TqS = a.Tq; d->TkS = a.Tk;
  d->cu_q = a.cu_q; d->cu_k = a.cu_k;
  if (a.cu_q) { d->q_bs = a.q_ts; d->o_bs = a.o_ts; }      // packed: "batch index" = first row of the sample (VARLEN_ROWS)
  if (a.cu_k) { d->k_bs = a.k_ts; d->v_bs = a.v_ts; }
  d->causal = a.causal;
  d->key_tokens = reinterpret_cast<const long long*>(a.key_tokens); d->pad_idx = a.pad_idx;
  d->key_pad_mask = a.key_pad_mask;
  d->scale = a.scale;
  d->drop = a.drop;
}

template <int HD, int NT>
static int launch_fwd(const AttnDev& d, cudaStream_t s) {
  const size_t smem = AttnSmem<HD>::fwd_bytes(d.TQP, d.TKP);
  B200_REQUIRE(smem <= 227 * 1024, "attention fwd: %zu B of shared memory needed (> 227 KB)", smem);
  static size_t configured = 48 * 1024;      // the default limit: never lower it
  if (smem > configured) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HD, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  B200_CHECK_CUDA(launch_kernel(attn_fwd_kernel<HD, NT>, dim3(d.H, d.B), dim3(128), smem, s, true, 1, d));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}
template <int HD, int NT>
static int launch_bwd(const AttnDev& d, cudaStream_t s) {
  const size_t smem = AttnSmem<HD>::bwd_bytes(d.TQP, d.TKP);
  B200_REQUIRE(smem <= 227 * 1024, "attention bwd: %zu B of shared memory needed (> 227 KB)", smem);
  static size_t configured = 48 * 1024;      // the default limit: never lower it
  if (smem > configured) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<HD, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  B200_CHECK_CUDA(launch_kernel(attn_bwd_kernel<HD, NT>, dim3(d.H, d.B), dim3(128), smem, s, true, 1, d));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int HD, int NT, int MAXI, int MINB = 2, int NW = 8>
static int launch_bwd2(const AttnDev& d, cudaStream_t s) {
  constexpr int LD = HD + 8;
  const int LDP = d.TKP + 8;
  const size_t smem = static_cast<size_t>(2 * d.TQP + 2 * d.TKP) * LD * 2 + static_cast<size_t>(d.TQP) * (LDP > LD ? LDP : LD) * 2 +
                      (d.TKP + 2 * d.TQP) * sizeof(float) + 16;
  B200_REQUIRE(smem <= 227 * 1024, "attention bwd: %zu B of shared memory needed (> 227 KB)", smem);
  static size_t configured = 48 * 1024;      // the default limit: never lower it
  if (smem > configured) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel<HD, NT, MAXI, MINB, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  B200_CHECK_CUDA(launch_kernel(attn_bwd2_kernel<HD, NT, MAXI, MINB, NW>, dim3(d.H, d.B), dim3(NW * 32), smem, s, true, 1, d));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static constexpr int kBwdShortMinB = 4;   // CTAs per SM the short-key staged backward is compiled for; B200_ATTN_BWD_MINB overrides
// (long keys: 12 / 16 warps per CTA at 80 / 64 registers measured 188 / 178 us against 174 us for 8 warps at cfg2's
// cross attention -- two CTAs per SM by shared memory either way, and the phases are not warp-count bound)
// staged backward when its register-held work items fit (<= 8 warps x MAXI), else the two-phase one
template <int HD>
static int dispatch_bwd(const AttnDev& d, cudaStream_t s, int nt_fallback) {
  static const bool force_v1 = getenv("B200_ATTN_BWD_V1") != nullptr;
  const int n_qt = d.TQP / 16;
  if (!force_v1) {
    if (d.TKP > 96) {
      const int items = n_qt * ((d.TKP + 47) / 48);
      if (items <= 16) return launch_bwd2<HD, 6, 2>(d, s);
      if (items <= 24) return launch_bwd2<HD, 6, 3>(d, s);
    } else {
      const int items = n_qt * (d.TKP / 16);
      if (items <= 16 && HD <= 64) {
        // short keys (self attention of a caption): little shared memory, so the register cap decides how many
        // CTAs share an SM; 3 fit without spills (80 registers), 4 with a few spilled words
        static const int minb = getenv("B200_ATTN_BWD_MINB") ? atoi(getenv("B200_ATTN_BWD_MINB")) : kBwdShortMinB;
        if (minb == 3) return launch_bwd2<HD, 2, 2, 3>(d, s);
        if (minb == 4) return launch_bwd2<HD, 2, 2, 4>(d, s);
      }
      if (items <= 16) return launch_bwd2<HD, 2, 2>(d, s);
      if (items <= 24) return launch_bwd2<HD, 2, 3>(d, s);
      const int items48 = n_qt * ((d.TKP + 47) / 48);
      if (items48 <= 16) return launch_bwd2<HD, 6, 2>(d, s);
      if (items48 <= 24) return launch_bwd2<HD, 6, 3>(d, s);
    }
  }
  return nt_fallback == 8 ? launch_bwd<HD, 8>(d, s) : launch_bwd<HD, 4>(d, s);
}

// one function (and one `configured` high-water mark) per kernel instantiation: the dynamic shared-memory
// limit of a kernel must never be lowered below what a later, larger launch of the SAME kernel needs
template <int HD, int NT>
static int launch_fwd_split(const AttnDev& d, int KS, int items, size_t smem, cudaStream_t s) {
  auto kern = attn_fwd_split_kernel<HD, NT>;
  static size_t configured = 48 * 1024;      // the default limit
  if (smem > configured) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  B200_CHECK_CUDA(launch_kernel(kern, dim3(d.H, d.B), dim3(items * 32), smem, s, true, 1, d, KS));
  note_launch();
  return 0;
}

// split-key forward: KS key ranges per query tile, one warp per (query tile, range).  Returns 1 when
// the shape does not fit (too many warps, ranges longer than NT*8 keys, partial tiles larger than the
// K/V space they alias) and the caller falls back to the query-tile-per-warp kernel.
template <int HD>
static int try_launch_fwd_split(const AttnDev& d, cudaStream_t s) {
  static const bool off = getenv("B200_ATTN_FWD_V1") != nullptr;
  if (off) return 1;
  constexpr int LD = HD + 8, LDO = HD + 4;
  const int n_qt = d.TQP / 16, n_kt = d.TKP / 16;
  if (n_qt > 16) return 1;
  int KS = 16 / n_qt;
  if (KS > n_kt) KS = n_kt;
  // partial O tiles alias the K/V tiles
  while (KS > 1 && static_cast<size_t>(n_qt) * KS * 16 * LDO * 4 > static_cast<size_t>(2) * d.TKP * LD * 2) --KS;
  const int tiles_per = (n_kt + KS - 1) / KS;            // longest range, in 16-key tiles
  if (tiles_per > 4) return 1;
  const int items = n_qt * KS;
  const size_t smem = static_cast<size_t>(d.TQP + 2 * d.TKP) * LD * 2 + (d.TKP + items * 32) * sizeof(float) + 16;
  if (smem > 227 * 1024) return 1;
  if (tiles_per <= 2) return launch_fwd_split<HD, 4>(d, KS, items, smem, s);
  return launch_fwd_split<HD, 8>(d, KS, items, smem, s);
}

int attn_fwd(const AttnArgs& a, cudaStream_t s) {
  if (int rc = check_common(a)) return rc;
  if (attn_tc_supported(a)) return attn_tc_fwd(a, s);      // tcgen05 / TMEM path (attention_tc.cu)
  AttnDev d;
  fill_dev(a, &d);
  {
    int rc = 1;
    switch (a.hd) {
      case 32: rc = try_launch_fwd_split<32>(d, s); break;
      case 64: rc = try_launch_fwd_split<64>(d, s); break;
      case 96: rc = try_launch_fwd_split<96>(d, s); break;
      default: rc = try_launch_fwd_split<128>(d, s); break;
    }
    if (rc <= 0) return rc;
  }
  switch (a.hd) {
    case 32: return launch_fwd<32, 8>(d, s);
    case 64: return launch_fwd<64, 8>(d, s);
    case 96: return launch_fwd<96, 8>(d, s);
    default: return launch_fwd<128, 8>(d, s);
  }
}

int attn_bwd(const AttnArgs& a, const AttnGrads& gr, cudaStream_t s) {
  if (int rc = check_common(a)) return rc;
  B200_REQUIRE(a.lse != nullptr, "attention bwd: lse is required");
  auto ok = [](const void* ptr, long long bs, long long ts) {
    return ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && bs % 8 == 0 && ts % 8 == 0;
  };
  B200_REQUIRE(ok(gr.d_o, gr.do_bs, gr.do_ts) && ok(gr.dq, gr.dq_bs, gr.dq_ts) && ok(gr.dk, gr.dk_bs, gr.dk_ts) && ok(gr.dv, gr.dv_bs, gr.dv_ts),
               "attention bwd: dO/dQ/dK/dV must be non-null, 16-byte aligned, strides multiples of 8");
  if (attn_tc_bwd_supported(a, gr)) return attn_tc_bwd(a, gr, s);   // tcgen05 / TMEM path (attention_tc.cu)
  AttnDev d;
  fill_dev(a, &d);
  d.d_o = gr.d_o; d.do_bs = gr.do_bs; d.do_ts = gr.do_ts;
  d.dq = gr.dq; d.dq_bs = gr.dq_bs; d.dq_ts = gr.dq_ts;
  d.dk = gr.dk; d.dk_bs = gr.dk_bs; d.dk_ts = gr.dk_ts;
  d.dv = gr.dv; d.dv_bs = gr.dv_bs; d.dv_ts = gr.dv_ts;
  if (a.cu_q) { d.do_bs = gr.do_ts; d.dq_bs = gr.dq_ts; }
  if (a.cu_k) { d.dk_bs = gr.dk_ts; d.dv_bs = gr.dv_ts; }
  switch (a.hd) {
    case 32: return dispatch_bwd<32>(d, s, 8);
    case 64: return dispatch_bwd<64>(d, s, 8);
    case 96: return dispatch_bwd<96>(d, s, 4);
    default: return dispatch_bwd<128>(d, s, 4);
  }
}

}  // namespace b200
