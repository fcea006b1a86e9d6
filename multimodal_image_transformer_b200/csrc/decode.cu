// Memory-bound kernels of KV-cached generation (replaces the O(T^2) full-prefix recompute of the
// reference greedy loop, model.py:221-240, where every step re-runs the whole decoder and
// re-projects the image memory in every layer).  Per step and layer the dominant traffic is the
// per-image cross-attention K/V (S x E x 2 bf16), streamed exactly once with coalesced 128-bit
// loads from a head-major layout; all beams of an image share that stream.
#include "decode.cuh"
#include <math.h>
#include <stdlib.h>

namespace b200 {

static inline int cdiv(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------
__global__ void kv_to_head_major_kernel(const bf16* __restrict__ kv, bf16* __restrict__ k_hm,
                                        bf16* __restrict__ v_hm, int B, int S, int H, int hd) {
  pdl_wait();
  pdl_trigger();
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
  B200_CHECK_CUDA(launch_kernel(kv_to_head_major_kernel, dim3(cdiv(total, 256)), dim3(256), 0, s, true, 1, kv, k_hm, v_hm, B, S, H, hd));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// Single-position cross attention over a K/V stream that lives in HBM (flash-decoding, one query
// per hypothesis).  Work item = one (group, head): a contiguous [nkeys, HD] K plane and V plane.
// Persistent CTAs; a producer warp streams 64-key chunks of K and V with 1-D bulk copies
// (cp.async.bulk + mbarrier) through a shared-memory ring that runs ahead ACROSS work items, so the
// HBM pipe never drains between heads; four consumer warps keep an online-softmax state per lane
// group (GL lanes share one key row, 16 bytes each).  A lane group takes its keys of a chunk four
// at a time: four independent score chains, ONE running-max update, four accumulations -- the
// loop-carried dependency is per batch, not per key.  The 4*KPW partial states are merged once per
// item; the next item's queries are fetched before the merge.  All NQ queries of a group (the
// beams of one image) share the stream.
// ------------------------------------------------------------------------------------------
struct DecAttnArgs {
  const bf16* q; long long q_rs;
  const bf16* k; const bf16* v;     // [groups][H][kv_len][HD]
  int kv_len, nkeys;                // row capacity / valid keys
  bf16* o; long long o_rs;
  int groups, nq, H;
  const unsigned char* key_pad;     // [groups][nkeys] or null
  float scale;
  // optional L2 warm-up for the NEXT cross-attention launch (the following layer's K / V planes): the
  // producer warp, once its own stream is issued, requests the first pf_bytes of each plane into L2
  // (cp.async.bulk.prefetch.L2).  Hint only: results do not depend on it.  Measured on B200: no gain at any
  // size (the requests queue in front of the launch chain's own weight loads), so the engine leaves it off.
  const bf16* pf_k; const bf16* pf_v; long long pf_bytes;
  int stream_evict_first;           // 1: the K/V stream is loaded with the L2 evict-first priority (read once)
  int tail16;                       // 1: a last chunk of < 64 keys is fetched as 16-row boxes
  // optional dynamic work distribution: sched[0] = next (image, head) item, sched[1] = units that ran dry; both
  // zero before the launch and zero again after it (the last unit resets them).  A unit that starts late
  // (its SM was busy with another stream's kernel) then simply takes fewer items.
  int* sched;
};

static constexpr int DEC_CK = 64;       // keys per chunk
static constexpr int DEC_STAGES = 3;
static constexpr int DEC_THREADS = 160; // 4 consumer warps + 1 producer warp

__device__ __forceinline__ float quad_max_dec(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum_dec(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ float exp_diff(float a, float m) {   // exp(a - m) with exp(-inf - -inf) = 0
  return (a == -INFINITY) ? 0.f : __expf(a - m);
}
__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

template <int HD, int NQ>
__global__ void __launch_bounds__(DEC_THREADS)
attn_decode_stream_kernel(const DecAttnArgs a) {
  constexpr int GL = (HD <= 32) ? 4 : (HD <= 64 ? 8 : 16);   // lanes cooperating on one key row
  constexpr int KPW = 32 / GL;                               // keys per warp per iteration
  constexpr int KPI = 4 * KPW;                               // keys per CTA per iteration
  constexpr int NIT = DEC_CK / KPI;                          // iterations per chunk (multiple of 4)
  constexpr int PLANE = DEC_CK * HD * 2;                     // bytes of one K (or V) chunk
  constexpr int STAGE = 2 * PLANE;
  constexpr int RED_STRIDE = HD + 2;
  extern __shared__ __align__(128) uint8_t sm_dec[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm_dec + DEC_STAGES * STAGE);
  uint64_t* empty = full + DEC_STAGES;
  float* red = reinterpret_cast<float*>(empty + DEC_STAGES + 2);   // [2][4][NQ][RED_STRIDE]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEC_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 4); }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  const int n_items = a.groups * a.H;
  const int nch = (a.nkeys + DEC_CK - 1) / DEC_CK;
  const long long plane_elems = static_cast<long long>(a.kv_len) * HD;

  if (warp == 4) {
    // ============================ producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const bf16* kp = a.k + item * plane_elems;
        const bf16* vp = a.v + item * plane_elems;
        for (int c = 0; c < nch; ++c) {
          const int nk = min(DEC_CK, a.nkeys - c * DEC_CK);
          const uint32_t bytes = static_cast<uint32_t>(nk) * HD * 2;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* dst = sm_dec + stage * STAGE;
          mbar_arrive_expect_tx(&full[stage], 2 * bytes);
          bulk_load_1d(dst, kp + static_cast<long long>(c) * DEC_CK * HD, bytes, &full[stage]);
          bulk_load_1d(dst + PLANE, vp + static_cast<long long>(c) * DEC_CK * HD, bytes, &full[stage]);
          if (++stage == DEC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ============================== consumers ==============================
  const int gl = lane % GL, kg = lane / GL;
  const int d0 = gl * 8;
  const bool active = d0 < HD;
  int stage = 0, par = 0;
  uint32_t phase = 0;

  uint4 qraw[NQ];
  auto fetch_q = [&](int item) {
    const int g = item / a.H, h = item % a.H;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      qraw[i] = make_uint4(0u, 0u, 0u, 0u);
      if (active && i < a.nq && item < n_items)
        qraw[i] = *reinterpret_cast<const uint4*>(a.q + (static_cast<long long>(g) * a.nq + i) * a.q_rs + h * HD + d0);
    }
  };
  fetch_q(blockIdx.x);

  for (int item = blockIdx.x; item < n_items; item += gridDim.x, par ^= 1) {
    const int g = item / a.H, h = item % a.H;
    float qr[NQ][8], m[NQ], l[NQ], acc[NQ][8];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      m[i] = -INFINITY;
      l[i] = 0.f;
      unpack8f(qraw[i], qr[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) { qr[i][e] *= a.scale; acc[i][e] = 0.f; }
    }

    for (int c = 0; c < nch; ++c) {
      mbar_wait(&full[stage], phase);
      const uint8_t* sk = sm_dec + stage * STAGE;
      const uint8_t* sv = sk + PLANE;
      const int nk = min(DEC_CK, a.nkeys - c * DEC_CK);
      const unsigned char* padp = a.key_pad ? a.key_pad + static_cast<long long>(g) * a.nkeys + c * DEC_CK : nullptr;
#pragma unroll
      for (int it0 = 0; it0 < NIT; it0 += 4) {
        if (it0 * KPI + warp * KPW >= nk) break;            // warp-uniform: nothing left for this warp
        float sc[4][NQ];
        uint4 vu[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = (it0 + u) * KPI + warp * KPW + kg;
          bool valid = j < nk;
          uint4 ku = make_uint4(0u, 0u, 0u, 0u);
          vu[u] = make_uint4(0u, 0u, 0u, 0u);
          if (valid && active) {
            ku = *reinterpret_cast<const uint4*>(sk + j * (HD * 2) + d0 * 2);
            vu[u] = *reinterpret_cast<const uint4*>(sv + j * (HD * 2) + d0 * 2);
          }
          if (valid && padp != nullptr && padp[j] != 0) valid = false;
          float kf[8];
          unpack8f(ku, kf);
#pragma unroll
          for (int i = 0; i < NQ; ++i) {
            float part = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) part = fmaf(qr[i][e], kf[e], part);
#pragma unroll
            for (int off = GL / 2; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
            sc[u][i] = valid ? part : -INFINITY;
          }
        }
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          const float cm = fmaxf(fmaxf(sc[0][i], sc[1][i]), fmaxf(sc[2][i], sc[3][i]));
          if (cm > m[i]) {                       // uniform within the lane group
            const float corr = exp_diff(m[i], cm);
            l[i] *= corr;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[i][e] *= corr;
            m[i] = cm;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float vf[8];
          unpack8f(vu[u], vf);
#pragma unroll
          for (int i = 0; i < NQ; ++i) {
            const float p = exp_diff(sc[u][i], m[i]);
            l[i] += p;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(p, vf[e], acc[i][e]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == DEC_STAGES) { stage = 0; phase ^= 1; }
    }
    fetch_q(item + gridDim.x);   // next item's queries travel while this item is merged

    // merge the KPW lane groups of this warp
#pragma unroll
    for (int off = GL; off < 32; off <<= 1) {
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const float mo = __shfl_xor_sync(0xffffffffu, m[i], off);
        const float lo = __shfl_xor_sync(0xffffffffu, l[i], off);
        const float mm = fmaxf(m[i], mo);
        const float cs = exp_diff(m[i], mm), co = exp_diff(mo, mm);
        l[i] = l[i] * cs + lo * co;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float ao = __shfl_xor_sync(0xffffffffu, acc[i][e], off);
          acc[i][e] = acc[i][e] * cs + ao * co;
        }
        m[i] = mm;
      }
    }
    float* rp = red + (par * 4 + warp) * NQ * RED_STRIDE;
    if (kg == 0 && active) {
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
#pragma unroll
        for (int e = 0; e < 8; ++e) rp[i * RED_STRIDE + d0 + e] = acc[i][e];
        if (gl == 0) { rp[i * RED_STRIDE + HD] = m[i]; rp[i * RED_STRIDE + HD + 1] = l[i]; }
      }
    }
    named_barrier_sync(1, 128);
    const float* rb = red + par * 4 * NQ * RED_STRIDE;
    for (int idx = threadIdx.x; idx < NQ * HD; idx += 128) {
      const int i = idx / HD, d = idx % HD;
      if (i < a.nq) {
        float mm = -INFINITY;
#pragma unroll
        for (int w = 0; w < 4; ++w) mm = fmaxf(mm, rb[(w * NQ + i) * RED_STRIDE + HD]);
        float ls = 0.f, val = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float cw = exp_diff(rb[(w * NQ + i) * RED_STRIDE + HD], mm);
          ls = fmaf(rb[(w * NQ + i) * RED_STRIDE + HD + 1], cw, ls);
          val = fmaf(rb[(w * NQ + i) * RED_STRIDE + d], cw, val);
        }
        a.o[(static_cast<long long>(g) * a.nq + i) * a.o_rs + h * HD + d] = __float2bfloat16(ls > 0.f ? val / ls : 0.f);
      }
    }
  }
}

template <int HD, int NQ>
static int launch_attn_decode(const DecAttnArgs& a, cudaStream_t s) {
  auto kern = attn_decode_stream_kernel<HD, NQ>;
  const size_t smem = static_cast<size_t>(DEC_STAGES) * 2 * DEC_CK * HD * 2 + (2 * DEC_STAGES + 2) * sizeof(uint64_t) +
                      static_cast<size_t>(2) * 4 * NQ * (HD + 2) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  int per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
  const int reg_cap = NQ > 1 ? 3 : 4;       // 160 threads x ~100 (NQ = 1) / 128 (NQ = 4) registers
  per_sm = per_sm < 1 ? 1 : (per_sm > reg_cap ? reg_cap : per_sm);
  const int items = a.groups * a.H;
  const int cap = device_sm_count() * per_sm;
  B200_CHECK_CUDA(launch_kernel(kern, dim3(items < cap ? items : cap), dim3(DEC_THREADS), smem, s, true, 1, a));
  note_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Tensor-core variant of the stream kernel (HD = 64 / 128, nq <= 8): the scalar kernel above is
// issue-bound (64 % issue-active at 55 % of DRAM peak: ~36 warp-instructions per key), so here the
// 64-key chunks arrive as 128-byte-swizzled TMA tiles and each consumer warp runs S = Q K^T and
// P V for its 16 keys of the chunk on mma.sync.m16n8k16 (Q = a 16-row tile whose first nq rows are
// the hypotheses of the image, the rest zero): ~3.5 warp-instructions per key, all beams for free.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dec_ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void dec_ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void dec_mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float dec_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

static constexpr int DECM_NQ = 4;       // rows of the partial-state exchange (nq <= 4)

// bytes of one producer/consumer unit: ring + barriers + merge buffer, rounded to the swizzle alignment
template <int HD, int DECM_STAGES, int NQR>
__host__ __device__ constexpr size_t decm_unit_bytes() {
  return ((static_cast<size_t>(DECM_STAGES) * 2 * (HD / 64) * DEC_CK * 128 + (2 * DECM_STAGES + 2) * sizeof(uint64_t) +
           static_cast<size_t>(2) * 4 * NQR * (HD + 2) * sizeof(float) + (DECM_STAGES + 4) * sizeof(int)) + 1023) / 1024 * 1024;
}

// SUB = 1: the CTA is one unit (1 producer + 4 consumer warps), MINB CTAs per SM.
// SUB > 1: a "fat" CTA of SUB independent units with private rings / barriers (one CTA per SM): a grid
//          smaller than the SM count then leaves whole SMs to concurrently running kernels -- the launch
//          chains of the other image partitions -- which a grid of small CTAs (spread over every SM by
//          the block scheduler) cannot do.
// NQR = query rows the partial-state exchange buffer holds (>= nq): 1 for greedy frees 6 KB per unit, which is
// what lets four units with a 3-deep ring share one SM.
template <int HD, int DECM_STAGES, int MINB, int SUB, int NQR>
__global__ void __launch_bounds__(DEC_THREADS * SUB, SUB == 1 ? MINB : 1)
attn_decode_mma_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                       const __grid_constant__ CUtensorMap tmap_k16, const __grid_constant__ CUtensorMap tmap_v16,
                       const DecAttnArgs a) {
  constexpr int NBOX = HD / 64;                       // 64-column TMA boxes per K (or V) chunk
  constexpr int BOX = DEC_CK * 128;                   // 8 KB: 64 rows x 128 B, 128-byte swizzled
  constexpr int STAGE = 2 * NBOX * BOX;
  constexpr int RED_STRIDE = HD + 2;
  extern __shared__ __align__(1024) uint8_t sm_decm_all[];
  const int sub = (SUB == 1) ? 0 : static_cast<int>(threadIdx.x) / DEC_THREADS;       // warp-uniform
  const int ltid = (SUB == 1) ? static_cast<int>(threadIdx.x) : static_cast<int>(threadIdx.x) % DEC_THREADS;
  const int vcta = static_cast<int>(blockIdx.x) * SUB + sub;                           // unit index / count
  const int vgrid = static_cast<int>(gridDim.x) * SUB;
  uint8_t* sm_decm = sm_decm_all + static_cast<size_t>(sub) * decm_unit_bytes<HD, DECM_STAGES, NQR>();
  uint64_t* full = reinterpret_cast<uint64_t*>(sm_decm + DECM_STAGES * STAGE);
  uint64_t* empty = full + DECM_STAGES;
  float* red = reinterpret_cast<float*>(empty + DECM_STAGES + 2);   // [2][4][NQR][RED_STRIDE]
  volatile int* item_ring = reinterpret_cast<volatile int*>(red + 2 * 4 * NQR * RED_STRIDE);   // [DECM_STAGES]: item whose first chunk is in the stage
  const bool dyn = a.sched != nullptr;

  const int warp = ltid >> 5, lane = ltid & 31;
  if (ltid == 0) {
    if ((smem_u32(sm_decm) & 1023u) != 0u) __trap();
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    if (a.tail16) { tma_prefetch_desc(&tmap_k16); tma_prefetch_desc(&tmap_v16); }
    for (int i = 0; i < DECM_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 4); }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  const int n_items = a.groups * a.H;
  const int nch = (a.nkeys + DEC_CK - 1) / DEC_CK;

  if (warp == 4) {
    // ============================ producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol = l2_policy_evict_first();
      const bool hint = a.stream_evict_first != 0;
      auto ld = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
        if (hint) tma_load_2d_hint(dst, map, bar, c0, c1, pol);
        else tma_load_2d(dst, map, bar, c0, c1);
      };
      // a short last chunk travels as 16-row boxes (same swizzled placement as the 64-row box: 2 KB per
      // 16 rows) instead of dragging up to 48 rows of the next item through L2 -> SM
      const int tail_tiles = (a.nkeys - (nch - 1) * DEC_CK + 15) / 16;
      const bool tail16 = a.tail16 != 0 && tail_tiles < DEC_CK / 16;
      for (int item = dyn ? atomicAdd(a.sched, 1) : vcta; item < n_items; item = dyn ? atomicAdd(a.sched, 1) : item + vgrid) {
        const int row0 = item * a.kv_len;
        for (int c = 0; c < nch; ++c) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* dst = sm_decm + stage * STAGE;
          if (dyn && c == 0) item_ring[stage] = item;      // published by the barrier arrive below (release)
          if (tail16 && c == nch - 1) {
            mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(tail_tiles) * 2u * NBOX * 2048u);
#pragma unroll
            for (int bx = 0; bx < NBOX; ++bx) {
              for (int tt = 0; tt < tail_tiles; ++tt) {
                ld(dst + bx * BOX + tt * 2048, &tmap_k16, &full[stage], bx * 64, row0 + c * DEC_CK + tt * 16);
                ld(dst + (NBOX + bx) * BOX + tt * 2048, &tmap_v16, &full[stage], bx * 64, row0 + c * DEC_CK + tt * 16);
              }
            }
          } else {
            mbar_arrive_expect_tx(&full[stage], STAGE);
#pragma unroll
            for (int bx = 0; bx < NBOX; ++bx) {
              ld(dst + bx * BOX, &tmap_k, &full[stage], bx * 64, row0 + c * DEC_CK);
              ld(dst + (NBOX + bx) * BOX, &tmap_v, &full[stage], bx * 64, row0 + c * DEC_CK);
            }
          }
          if (++stage == DECM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (dyn) {
        // no more items: tell the consumers (an empty stage carrying item -1), then account for this unit;
        // the last unit to get here re-arms the counters for the next launch
        mbar_wait(&empty[stage], phase ^ 1);
        item_ring[stage] = -1;
        mbar_arrive(&full[stage]);
        __threadfence();
        if (atomicAdd(a.sched + 1, 1) == vgrid - 1) {
          a.sched[0] = 0;
          a.sched[1] = 0;
          __threadfence();
        }
      }
    }
    __syncwarp();
    if (a.pf_bytes > 0) {
      // 4 KB requests, spread over the producer lanes of all CTAs
      constexpr long long CH = 4096;
      const long long nch_pf = (a.pf_bytes + CH - 1) / CH;
      for (long long c = static_cast<long long>(vcta) * 32 + lane; c < nch_pf; c += static_cast<long long>(vgrid) * 32) {
        const long long off = c * CH;
        const uint32_t n = static_cast<uint32_t>((a.pf_bytes - off < CH ? a.pf_bytes - off : CH) & ~15ll);
        if (n == 0) continue;
        l2_prefetch_bulk(reinterpret_cast<const uint8_t*>(a.pf_k) + off, n);
        l2_prefetch_bulk(reinterpret_cast<const uint8_t*>(a.pf_v) + off, n);
      }
    }
    return;
  }

  // ============================== consumers ==============================
  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;
  const float sl2 = a.scale * 1.4426950408889634f;
  int stage = 0, par = 0;
  uint32_t phase = 0;
  uint32_t qn[HD / 16][2];
  auto fetch_q = [&](int item) {
    const int grp = item / a.H, h = item % a.H;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      qn[ks][0] = qn[ks][1] = 0u;
      if (g < a.nq && item < n_items) {
        const bf16* qp = a.q + (static_cast<long long>(grp) * a.nq + g) * a.q_rs + h * HD + ks * 16 + 2 * t;
        qn[ks][0] = *reinterpret_cast<const uint32_t*>(qp);
        qn[ks][1] = *reinterpret_cast<const uint32_t*>(qp + 8);
      }
    }
  };
  if (!dyn) fetch_q(vcta);
  for (int item = vcta;; item += vgrid, par ^= 1) {
    if (dyn) {
      mbar_wait(&full[stage], phase);      // the item's first chunk (waited for again below: already complete)
      item = item_ring[stage];
      if (item < 0) break;
      fetch_q(item);
    } else if (item >= n_items) {
      break;
    }
    const int grp = item / a.H, h = item % a.H;
    // Q as mma A fragments: row g (< nq) = hypothesis g of this image, rows >= nq are zero
    uint32_t qa[HD / 16][4];
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      qa[ks][0] = qn[ks][0]; qa[ks][2] = qn[ks][1];
      qa[ks][1] = qa[ks][3] = 0u;
    }
    if (!dyn) fetch_q(item + vgrid);       // static stride: the next item's query travels while this item is processed
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }

    for (int c = 0; c < nch; ++c) {
      mbar_wait(&full[stage], phase);
      const int nk = min(DEC_CK, a.nkeys - c * DEC_CK);
      const int kw0 = warp * 16;                       // this warp's keys of the chunk
      if (kw0 < nk) {                                  // warp-uniform
        const uint32_t sk = smem_u32(sm_decm + stage * STAGE);
        const uint32_t sv = sk + NBOX * BOX;
        float sc[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          uint32_t b[4];
          const int row = kw0 + (mi >> 1) * 8 + r8;
          const int unit = (ks & 3) * 2 + (mi & 1);
          dec_ldsm_x4(b, sk + (ks >> 2) * BOX + row * 128 + ((unit ^ r8) << 4));
          dec_mma16816(sc[0], qa[ks], b[0], b[1]);
          dec_mma16816(sc[1], qa[ks], b[2], b[3]);
        }
        const unsigned char* padp = a.key_pad ? a.key_pad + static_cast<long long>(grp) * a.nkeys + c * DEC_CK : nullptr;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = kw0 + j * 8 + 2 * t + (e & 1);
            bool valid = key < nk;
            if (valid && padp != nullptr && padp[key] != 0) valid = false;
            const float v = valid ? sc[j][e] * sl2 : -INFINITY;
            sc[j][e] = v;
            mx[e >> 1] = fmaxf(mx[e >> 1], v);
          }
        }
        float alpha[2], ms[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float mn = fmaxf(m_run[r], quad_max_dec(mx[r]));
          alpha[r] = (m_run[r] == -INFINITY) ? 0.f : dec_ex2(m_run[r] - mn);
          m_run[r] = mn;
          ms[r] = (mn == -INFINITY) ? 0.f : mn;
        }
        float ps[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float pv = dec_ex2(sc[j][e] - ms[e >> 1]);
            sc[j][e] = pv;
            ps[e >> 1] += pv;
          }
        }
        l_run[0] = l_run[0] * alpha[0] + ps[0];
        l_run[1] = l_run[1] * alpha[1] + ps[1];
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
          o[i][0] *= alpha[0]; o[i][1] *= alpha[0]; o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
        }
        uint32_t pa[4];
        pa[0] = pack_bf16(sc[0][0], sc[0][1]); pa[1] = pack_bf16(sc[0][2], sc[0][3]);
        pa[2] = pack_bf16(sc[1][0], sc[1][1]); pa[3] = pack_bf16(sc[1][2], sc[1][3]);
#pragma unroll
        for (int dn = 0; dn < HD / 8; dn += 2) {
          uint32_t b[4];
          const int row = kw0 + (mi & 1) * 8 + r8;
          const int unit = (dn & 7) + (mi >> 1);
          dec_ldsm_x4_t(b, sv + (dn >> 3) * BOX + row * 128 + ((unit ^ r8) << 4));
          dec_mma16816(o[dn], pa, b[0], b[1]);
          dec_mma16816(o[dn + 1], pa, b[2], b[3]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == DECM_STAGES) { stage = 0; phase ^= 1; }
    }

    // ---- merge the four warps' states of rows < nq (row g of the tile lives in lanes 4g .. 4g+3)
    const float lsum = quad_sum_dec(l_run[0]);
    float* rp = red + (par * 4 + warp) * NQR * RED_STRIDE;
    if (g < a.nq) {
#pragma unroll
      for (int i = 0; i < HD / 8; ++i)
        *reinterpret_cast<float2*>(rp + g * RED_STRIDE + i * 8 + 2 * t) = make_float2(o[i][0], o[i][1]);
      if (t == 0) { rp[g * RED_STRIDE + HD] = m_run[0]; rp[g * RED_STRIDE + HD + 1] = lsum; }
    }
    named_barrier_sync(1 + sub, 128);
    const float* rb = red + par * 4 * NQR * RED_STRIDE;
    for (int idx = ltid; idx < a.nq * HD; idx += 128) {
      const int i = idx / HD, d = idx % HD;
      float mm = -INFINITY;
#pragma unroll
      for (int w = 0; w < 4; ++w) mm = fmaxf(mm, rb[(w * NQR + i) * RED_STRIDE + HD]);
      float ls = 0.f, val = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float mw = rb[(w * NQR + i) * RED_STRIDE + HD];
        const float cw = (mw == -INFINITY) ? 0.f : dec_ex2(mw - mm);
        ls = fmaf(rb[(w * NQR + i) * RED_STRIDE + HD + 1], cw, ls);
        val = fmaf(rb[(w * NQR + i) * RED_STRIDE + d], cw, val);
      }
      a.o[(static_cast<long long>(grp) * a.nq + i) * a.o_rs + h * HD + d] = __float2bfloat16(ls > 0.f ? val / ls : 0.f);
    }
  }
}

template <int HD, int DECM_STAGES, int MINB, int SUB, int NQR>
static int launch_attn_decode_mma(const DecAttnArgs& a, int grid_cap, cudaStream_t s) {
  B200_REQUIRE(a.nq <= NQR, "attn_decode: %d query rows for an exchange buffer of %d", a.nq, NQR);
  auto kern = attn_decode_mma_kernel<HD, DECM_STAGES, MINB, SUB, NQR>;
  const size_t unit = decm_unit_bytes<HD, DECM_STAGES, NQR>();
  const size_t smem = unit * SUB + 1024;
  B200_REQUIRE(smem <= 227 * 1024, "attn_decode: %zu B of shared memory for %d units per CTA", smem, SUB);
  static bool configured = false;
  if (!configured) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  const long long rows = static_cast<long long>(a.groups) * a.H * a.kv_len;
  CUtensorMap tk, tv;
  if (int rc = make_tmap_2d_bf16(&tk, a.k, HD, static_cast<uint64_t>(rows), static_cast<uint64_t>(HD) * 2, 64, DEC_CK)) return rc;
  if (int rc = make_tmap_2d_bf16(&tv, a.v, HD, static_cast<uint64_t>(rows), static_cast<uint64_t>(HD) * 2, 64, DEC_CK)) return rc;
  CUtensorMap tk16 = tk, tv16 = tv;
  if (a.tail16) {
    if (int rc = make_tmap_2d_bf16(&tk16, a.k, HD, static_cast<uint64_t>(rows), static_cast<uint64_t>(HD) * 2, 64, 16)) return rc;
    if (int rc = make_tmap_2d_bf16(&tv16, a.v, HD, static_cast<uint64_t>(rows), static_cast<uint64_t>(HD) * 2, 64, 16)) return rc;
  }
  const int items = a.groups * a.H;
  int cap;
  if (SUB == 1) {
    int per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > MINB ? MINB : per_sm);
    cap = device_sm_count() * per_sm;
  } else {
    cap = device_sm_count();
  }
  if (grid_cap > 0 && grid_cap < cap) cap = grid_cap;
  const int want = (items + SUB - 1) / SUB;
  B200_CHECK_CUDA(launch_kernel(kern, dim3(want < cap ? want : cap), dim3(DEC_THREADS * SUB), smem, s, true, 1, tk, tv, tk16, tv16, a));
  note_launch();
  return 0;
}

int attn_decode(const bf16* q, long long q_rs, const bf16* k, const bf16* v, int kv_len, int nkeys, bf16* o,
                long long o_rs, int groups, int nq, int H, int hd, const unsigned char* key_pad, float scale,
                cudaStream_t s, const bf16* pf_k, const bf16* pf_v, long long pf_bytes, int flags, int fat_grid, int* sched) {
  B200_REQUIRE(nq >= 1 && nq <= 4, "attn_decode: %d queries per group (supported: 1..4)", nq);
  B200_REQUIRE(nkeys >= 1 && nkeys <= kv_len, "attn_decode: nkeys %d outside [1, %d]", nkeys, kv_len);
  DecAttnArgs a = {};
  a.q = q; a.q_rs = q_rs; a.k = k; a.v = v; a.kv_len = kv_len; a.nkeys = nkeys; a.o = o; a.o_rs = o_rs;
  a.groups = groups; a.nq = nq; a.H = H; a.key_pad = key_pad; a.scale = scale;
  if (pf_k && pf_v && pf_bytes > 0 && (reinterpret_cast<uintptr_t>(pf_k) & 15) == 0 && (reinterpret_cast<uintptr_t>(pf_v) & 15) == 0) {
    a.pf_k = pf_k; a.pf_v = pf_v; a.pf_bytes = pf_bytes;
  }
  a.stream_evict_first = (flags & 1) ? 1 : 0;
  a.tail16 = (flags & 2) ? 1 : 0;
  a.sched = sched;
  static const bool scalar_only = getenv("B200_DEC_ATTN_SCALAR") != nullptr;     // A/B switch
  if (!scalar_only && nq <= DECM_NQ && (reinterpret_cast<uintptr_t>(k) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0) {
    // measured on B200 (512 images x 12 heads, S = 197): ring depth 2 with 4 CTAs per SM 0.82 ms per position,
    // depth 3 0.84, depth 4 (3 CTAs) 0.95, depth 1 0.92; 5-6 CTAs per SM (register cap 72 / 64, spills) 0.95-1.05
    // fat_grid > 0: one fat CTA per SM on at most fat_grid SMs (the other SMs stay free for concurrent kernels)
    const bool deep = (flags & 8) != 0;
    if (hd == 64) {
      if (fat_grid > 0 && nq == 1 && deep) return launch_attn_decode_mma<64, 3, 4, 4, 1>(a, fat_grid, s);   // 4 units x 3-deep ring
      return fat_grid > 0 ? launch_attn_decode_mma<64, 2, 4, 4, DECM_NQ>(a, fat_grid, s) : launch_attn_decode_mma<64, 2, 4, 1, DECM_NQ>(a, 0, s);
    }
    if (hd == 128) return fat_grid > 0 ? launch_attn_decode_mma<128, 2, 2, 2, DECM_NQ>(a, fat_grid, s) : launch_attn_decode_mma<128, 2, 2, 1, DECM_NQ>(a, 0, s);
  }
#define B200_AD(HDV)                                              \
  do {                                                            \
    if (nq == 1) return launch_attn_decode<HDV, 1>(a, s);         \
    return launch_attn_decode<HDV, 4>(a, s);                      \
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
// Self attention of one decode position with the cache append fused in.  One WARP per (row, head):
// the current position's q / k / v come straight from the packed QKV projection, k and v are
// written to cache row `pos` and attended together with the cached rows [0, pos).  GL lanes share a
// key row (16 bytes each), eight rows per lane group are in flight at a time; scores go through a
// per-warp shared-memory strip, so there is no block-level synchronisation at all.
// ------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128)
attn_decode_self_kernel(const bf16* __restrict__ qkv, long long rs, bf16* __restrict__ kcache,
                        bf16* __restrict__ vcache, int max_len, int pos, bf16* __restrict__ o, long long o_rs,
                        int n_items, int H, float scale, const int64_t* __restrict__ seq, long long seq_ld,
                        long long pad_idx) {
  constexpr int GL = (HD <= 32) ? 4 : (HD <= 64 ? 8 : 16);
  constexpr int KPW = 32 / GL;
  constexpr int UNR = 8;
  extern __shared__ float sm_self[];    // [4][pos + 1]
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + warp;
  if (item >= n_items) return;
  const int nkeys = pos + 1;
  float* sc = sm_self + warp * nkeys;
  const int r = item / H, h = item % H;
  const int E = H * HD;
  const int gl = lane % GL, kg = lane / GL;
  const int d0 = gl * 8;
  const bool active = d0 < HD;
  // key padding of the reference (decoder.py:162: tgt_key_padding_mask = tokens == pad_idx, rebuilt from the
  // whole prefix on every generate() step, model.py:224-228): a PAD id in the prefix is never attended to
  const int64_t* seq_r = seq ? seq + static_cast<long long>(r) * seq_ld : nullptr;
  const bf16* qp = qkv + static_cast<long long>(r) * rs + h * HD + d0;
  bf16* kb = kcache + static_cast<long long>(item) * max_len * HD + d0;
  bf16* vb = vcache + static_cast<long long>(item) * max_len * HD + d0;
  uint4 qu = make_uint4(0u, 0u, 0u, 0u), kn = qu, vn = qu;
  if (active) {
    qu = *reinterpret_cast<const uint4*>(qp);
    kn = *reinterpret_cast<const uint4*>(qp + E);
    vn = *reinterpret_cast<const uint4*>(qp + 2 * E);
    if (kg == 0) {
      *reinterpret_cast<uint4*>(kb + static_cast<long long>(pos) * HD) = kn;
      *reinterpret_cast<uint4*>(vb + static_cast<long long>(pos) * HD) = vn;
    }
  }
  float qf[8];
  unpack8f(qu, qf);
#pragma unroll
  for (int e = 0; e < 8; ++e) qf[e] *= scale;

  // ---- scores
  for (int j0 = 0; j0 < nkeys; j0 += KPW * UNR) {
    uint4 ku[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + u * KPW + kg;
      ku[u] = (j == pos) ? kn : make_uint4(0u, 0u, 0u, 0u);
      if (active && j < pos) ku[u] = ldg_nc_v4(kb + static_cast<long long>(j) * HD);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + u * KPW + kg;
      float kf[8];
      unpack8f(ku[u], kf);
      float part = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) part = fmaf(qf[e], kf[e], part);
#pragma unroll
      for (int off = GL / 2; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
      if (gl == 0 && j < nkeys) sc[j] = (seq_r && seq_r[j] == pad_idx) ? -INFINITY : part;
    }
  }
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < nkeys; j += 32) mx = fmaxf(mx, sc[j]);
  mx = warp_max(mx);
  if (mx == -INFINITY) mx = 0.f;      // every key is PAD (PAD at position 0): the row is 0, not NaN
  float sum = 0.f;
  for (int j = lane; j < nkeys; j += 32) {
    const float p = __expf(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  __syncwarp();

  // ---- o = P V
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j0 = 0; j0 < nkeys; j0 += KPW * UNR) {
    uint4 vu[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + u * KPW + kg;
      vu[u] = (j == pos) ? vn : make_uint4(0u, 0u, 0u, 0u);
      if (active && j < pos) vu[u] = ldg_nc_v4(vb + static_cast<long long>(j) * HD);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + u * KPW + kg;
      const float p = j < nkeys ? sc[j] : 0.f;
      float vf[8];
      unpack8f(vu[u], vf);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, vf[e], acc[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int off = GL; off < 32; off <<= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
  if (kg == 0 && active) {
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    uint4 out;
    out.x = pack_bf16(acc[0] * inv, acc[1] * inv); out.y = pack_bf16(acc[2] * inv, acc[3] * inv);
    out.z = pack_bf16(acc[4] * inv, acc[5] * inv); out.w = pack_bf16(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(o + static_cast<long long>(r) * o_rs + h * HD + d0) = out;
  }
}

int attn_decode_append(const bf16* qkv, long long qkv_rs, bf16* kcache, bf16* vcache, int max_len, int pos, bf16* o,
                       long long o_rs, int R, int H, int hd, float scale, cudaStream_t s, const int64_t* seq,
                       long long seq_ld, long long pad_idx) {
  B200_REQUIRE(pos >= 0 && pos < max_len, "attn_decode_append: position %d outside the cache (%d rows)", pos, max_len);
  B200_REQUIRE(max_len <= 4096, "attn_decode_append: cache rows %d > 4096", max_len);
  const int items = R * H;
  const size_t smem = static_cast<size_t>(4) * (pos + 1) * sizeof(float);
#define B200_AS(HDV)                                                                                              \
  B200_CHECK_CUDA(launch_kernel(attn_decode_self_kernel<HDV>, dim3(cdiv(items, 4)), dim3(128), smem, s, true, 1, \
                                qkv, qkv_rs, kcache, vcache, max_len, pos, o, o_rs, items, H, scale, seq, seq_ld, pad_idx))
  switch (hd) {
    case 32: B200_AS(32); break;
    case 64: B200_AS(64); break;
    case 96: B200_AS(96); break;
    case 128: B200_AS(128); break;
    default: B200_REQUIRE(false, "attn_decode_append: head dim %d not in {32,64,96,128}", hd);
  }
#undef B200_AS
  note_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------
__global__ void greedy_update_kernel(const int64_t* __restrict__ next_ids, int64_t* __restrict__ cur,
                                     int64_t* __restrict__ out_tokens, int* __restrict__ out_len,
                                     unsigned char* __restrict__ finished, int* __restrict__ n_finished, int R,
                                     int max_len, int pos, long long end_id, long long pad_id) {
  pdl_wait();
  pdl_trigger();
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
  B200_CHECK_CUDA(launch_kernel(greedy_update_kernel, dim3(cdiv(R, 128)), dim3(128), 0, s, true, 1, next_ids, cur_tokens, out_tokens, out_len, finished, n_finished,
                                                    R, max_len, pos, end_id, pad_id));
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
  pdl_wait();
  pdl_trigger();
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
  B200_CHECK_CUDA(launch_kernel(beam_topk_kernel, dim3(B), dim3(256), 0, s, true, 1, logits, beam_scores, finished, beam, V, end_id, first_step, out_tokens,
                                     out_parent, out_scores));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// sequences / flags / scores follow their parents; the chosen token is appended at pos+1
__global__ void beam_advance_kernel(const int64_t* __restrict__ seq_in, int64_t* __restrict__ seq_out,
                                    const unsigned char* __restrict__ fin_in, unsigned char* __restrict__ fin_out,
                                    const int64_t* __restrict__ tokens, const int* __restrict__ parent, int R,
                                    int beam, int max_len, int pos, long long end_id) {
  pdl_wait();
  pdl_trigger();
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
  B200_CHECK_CUDA(launch_kernel(beam_advance_kernel, dim3(R), dim3(64), 0, s, true, 1, seq_in, seq_out, fin_in, fin_out, tokens, parent, R, beam, max_len, pos, end_id));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// best hypothesis per image (first index on score ties), cut after its first END, PAD afterwards
__global__ void beam_finalize_kernel(const int64_t* __restrict__ seqs, const float* __restrict__ scores, int beam,
                                     int max_len, int n_tok, long long end_id, long long pad_id,
                                     int64_t* __restrict__ out_tokens, int* __restrict__ out_len,
                                     float* __restrict__ out_score) {
  pdl_wait();
  pdl_trigger();
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
  B200_CHECK_CUDA(launch_kernel(beam_finalize_kernel, dim3(B), dim3(32), 0, s, true, 1, seqs, scores, beam, max_len, n_tok, end_id, pad_id, out_tokens, out_len, out_score));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void fill_i64_kernel(int64_t* p, long long n, long long v) {
  pdl_wait();
  pdl_trigger();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}
int fill_i64(int64_t* p, long long n, long long v, cudaStream_t s) {
  B200_CHECK_CUDA(launch_kernel(fill_i64_kernel, dim3(cdiv(n, 256)), dim3(256), 0, s, true, 1, p, n, v));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// out[r*stride] = v for r < n   (column fill: START token of every sequence)
__global__ void fill_col_i64_kernel(int64_t* p, long long n, long long stride, long long v) {
  pdl_wait();
  pdl_trigger();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) p[i * stride] = v;
}
int fill_col_i64(int64_t* p, long long n, long long stride, long long v, cudaStream_t s) {
  B200_CHECK_CUDA(launch_kernel(fill_col_i64_kernel, dim3(cdiv(n, 256)), dim3(256), 0, s, true, 1, p, n, stride, v));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void store_col_i64_kernel(int64_t* p, const int64_t* __restrict__ src, long long n, long long stride, long long col) {
  pdl_wait();
  pdl_trigger();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) p[i * stride + col] = src[i];
}
int store_col_i64(int64_t* p, const int64_t* src, long long n, long long stride, long long col, cudaStream_t s) {
  B200_CHECK_CUDA(launch_kernel(store_col_i64_kernel, dim3(cdiv(n, 256)), dim3(256), 0, s, true, 1, p, src, n, stride, col));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void cache_reorder_kernel(const bf16* __restrict__ src_k, const bf16* __restrict__ src_v,
                                     bf16* __restrict__ dst_k, bf16* __restrict__ dst_v,
                                     const int* __restrict__ parent, int beam, int H, int hd, int max_len, int npos) {
  pdl_wait();
  pdl_trigger();
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
  B200_CHECK_CUDA(launch_kernel(cache_reorder_kernel, grid, dim3(256), 0, s, true, 1, src_k, src_v, dst_k, dst_v, parent, beam, H, hd, max_len, npos));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
