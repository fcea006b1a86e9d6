// tcgen05 / TMEM / TMA attention for the training step (head dim 64, <= 64 query positions per item):
// replaces F.scaled_dot_product_attention reached from torch/nn/functional.py:6682 (forward) and its autograd
// backward for the shapes the caption decoder trains on -- self attention over the caption (Tq = Tk <= 64, causal +
// key padding from the token ids, decoder.py:158-162) and cross attention over the 50-257 image tokens (no mask in the
// reference, model.py:158; an optional key-padding mask is supported).  Other shapes keep the mma.sync kernels of
// attention.cu.
//
// One persistent CTA per SM walks the (image, head) items; per item
//   TMA producer (warp 0)   Q [64 x 64], K and V [Tk16 x 64] as 128-byte-swizzled tiles into a ring of item slots;
//   MMA issuer   (warp 1)   S = Q K^T as TWO M = 64 tcgen05.mma groups, one per half of the keys, written to the SAME
//                           TMEM columns at lane offsets 0 and 16 ("interleaved" M = 64 accumulators: rows 16w..16w+15
//                           of an M = 64 tile live in lanes 32w + {0..15}; the second tile takes lanes 32w + {16..31}).
//                           Every warp of a softmax group therefore holds 16 query rows x BOTH key halves: 32 busy
//                           lanes for a 47-row problem, row statistics combined with one shuffle;
//   softmax      (2 x 4 warps, alternating items)  tcgen05.ld the scores (thread = (row, key half)), two-pass softmax
//                           in registers, P as bf16 into shared memory in the K-major 128-byte-swizzled layout the
//                           tensor core reads (it overwrites the item's K tile, dead once S is complete);
//   MMA issuer              O = P V  (A = P from shared memory, B = V as stored = MN-major) into 64 more TMEM columns;
//   softmax group           tcgen05.ld O, scale by 1 / row sum, 128-byte row stores; log-sum-exp for the backward.
// TMEM holds NTM (S, O) stages, shared memory NSLOT item slots: the loads of item i+2 and the S product of item i+1 run
// under the softmax of item i.
//
// The backward kernel follows the same plan per item (see attn_tc_bwd_kernel below): S and dP = dO V^T into TMEM,
// P / dS through shared memory in bf16, dV = P^T dO and dK = dS^T Q as 128-row key tiles, dQ = dS K.
#include "attention.cuh"
#include "attention_tc.cuh"
#include <math.h>
#include <stdlib.h>

namespace b200 {

namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int TC_HD = 64;
constexpr int TC_THREADS = 384;            // 12 warps, see the role table at the kernels (ptxas caps registers as for 512 threads above 384)
constexpr int TC_MAX_SLOTS = 4;
constexpr int TC_MAX_TM = 4;
constexpr int ROW_BYTES = TC_HD * 2;       // 128: one swizzle row
constexpr int Q_BYTES = 64 * ROW_BYTES;    // 8 KB
constexpr int SLAB_BYTES64 = 64 * ROW_BYTES;   // one 64-key slab of P: 64 query rows x 128 B

// Bounded mbarrier wait in SEVEN instructions.  These kernels wait in ~25 inlined places and run six different warp
// roles at once: their hot loops have to fit the 32 KB L1.5 instruction cache together (measured on the backward
// kernel: with the clock64 / printf time-out of common.cuh's mbar_wait inlined everywhere, trace stamps and unrolled
// role bodies the binary was 61 KB and stall_no_inst was 58 % of the non-waiting samples of the elementwise warps).
// A protocol bug still surfaces as a trap (after 2^24 polls: >= 0.3 s), never as a hung GPU.
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 n;\n\t"
      "mov.u32 n, 0;\n\t"
      "TC_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TC_WAIT_DONE;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 p, n, 16777216;\n\t"
      "@p bra TC_WAIT_LOOP;\n\t"
      "trap;\n\t"
      "TC_WAIT_DONE:\n\t"
      "}"
      :
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// explicit shared-space accesses: pointers computed as `smem + runtime offset` otherwise compile to GENERIC LD.E / ST.E
// (measured on the first version of the backward kernel: the generic accesses were its top stall, stall_lg / long scoreboard)
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
// global loads that stay where they are written (a prefetch an item ahead must not sink below the next barrier wait)
__device__ __forceinline__ uint4 ldg_v4_here(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float ldg_f1_here(const float* p) {
  float r;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// ---- TMEM loads: 32 lanes x N consecutive 32-bit columns (thread = lane)
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// tcgen05.wait::ld that also names the destination registers of the load it completes: the compiler sees a
// read-modify-write of those registers, so no use of them can be scheduled above the wait (the loads are asynchronous;
// a plain asm volatile wait orders only against other volatile statements)
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
// the same for two 16-register loads
__device__ __forceinline__ void tmem_ld_wait_dep16x2(uint32_t (&a)[16], uint32_t (&b)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                 "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]),
                 "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                 "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15])
               :
               : "memory");
}
// Sweep nblk blocks of 32 fp32 columns starting at taddr: f(values, first column) per block, compact runtime loop.
// (Measured: two blocks per wait -- 64 columns of straight-line code, 166 registers -- ran 30 % SLOWER than one.)
template <class F>
__device__ __forceinline__ void tmem_sweep2(uint32_t taddr, int nblk, F&& f) {
#pragma unroll 1
  for (int blk = 0; blk < nblk; ++blk) {
    uint32_t ra[32];
    tmem_ld_32x32(taddr + blk * 32, ra);
    tmem_ld_wait();
    f(ra, blk * 32);
  }
}

// visit the `n` (multiple of 8) fp32 columns starting at taddr in chunks of 32 / 16 / 8: f(values, first column)
template <class F>
__device__ __forceinline__ void tmem_sweep(uint32_t taddr, int n, F&& f) {
  int c = 0;
  for (; c + 32 <= n; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + c, r);
    tmem_ld_wait();
    f(r, c);
  }
  if (c + 16 <= n) {
    uint32_t r[16];
    tmem_ld_32x16(taddr + c, r);
    tmem_ld_wait();
    f(r, c);
    c += 16;
  }
  if (c + 8 <= n) {
    uint32_t r[8];
    tmem_ld_32x8(taddr + c, r);
    tmem_ld_wait();
    f(r, c);
  }
}

// ring position without a division per trip (code size: see tc_wait)
struct RingPos {
  int idx = 0;
  uint32_t ph = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++idx == n) { idx = 0; ph ^= 1u; }
  }
};
// (image, head) of the CTA's items: item = first + i * stride, b = item / H, h = item % H
struct ItemPos {
  int b, h, db, dh, H;
  __device__ __forceinline__ ItemPos(int first, int stride, int H_) : b(first / H_), h(first % H_), db(stride / H_), dh(stride % H_), H(H_) {}
  __device__ __forceinline__ void advance() {
    b += db; h += dh;
    if (h >= H) { h -= H; ++b; }
  }
};

struct TcDev {
  int B, H, Tq, Tk, Tk16, half, nslab;
  int n_items;
  int causal, has_bias;
  int nslot, ntm, n_o, stage_cols;        // item slots; S stages and O stages in TMEM
  int k_off, v_off, slot_bytes;          // forward: K offset inside a Q|K ring entry, -, entry pitch (1024-byte multiples)
  int nv, v_base, v_bytes, n_p, p_base, p_bytes;   // forward: V ring (entries, offset, pitch) and the P buffers (item i uses i % n_p)
  int bias_off, bar_off;                 // byte offsets from the (aligned) shared-memory base
  const long long* key_tokens; long long pad_idx;
  const unsigned char* key_pad_mask;
  const int32_t* cu_q;                   // packed query rows (attention.cuh): sample b owns rows [cu_q[b], cu_q[b+1]); Tq = the largest
  float scale, sl2;                      // softmax scale and scale * log2(e)
  bf16* o; long long o_bs, o_ts;
  float* lse;
  DropCfg drop;
  // backward only
  const bf16* o_in; const bf16* d_o; long long do_bs, do_ts;
  bf16 *dq, *dk, *dv; long long dq_bs, dq_ts, dk_bs, dk_ts, dv_bs, dv_ts;
  int do_off, p_off, ds_off;
  long long* trace;                      // bring-up instrument: [item][16] clock64 stamps of CTA 0 (null = off)
};

// stamps of CTA 0's first 32 items: 0 producer issue, 1 S issued, 2 p_full seen, 3 PV issued, 4 s_full seen (softmax),
// 5 pass 1 done, 6 pass 2 done (P published), 7 o_full seen, 8 epilogue done
template <bool TRACE>
__device__ __forceinline__ void tc_stamp(const TcDev& p, int i, int slot) {
  if constexpr (TRACE) {
    if (blockIdx.x == 0 && i < 32) p.trace[i * 16 + slot] = clock64();
  }
}

__device__ __forceinline__ uint32_t tc_drop_pair(const TcDev& p, int bh, int row, int key) {
  return (static_cast<uint32_t>(bh) * p.Tq + row) * static_cast<uint32_t>((p.Tk + 1) >> 1) + (static_cast<uint32_t>(key) >> 1);
}

// byte offset of the 16-byte unit holding keys [key, key + 8) (key % 8 == 0) of query row `row` in the P / dS buffer:
// 64-key slabs of [64 rows][128 B], 128-byte swizzle (16-byte unit index XOR row % 8) -- as a K-major operand it is
// A[M = rows][K = keys] (P V, dS K), as an MN-major operand A[M = keys][K = rows] (P^T dO, dS^T Q)
__device__ __forceinline__ uint32_t p_unit_off(int row, int key) {
  return static_cast<uint32_t>((key >> 6) * SLAB_BYTES64 + row * ROW_BYTES + ((((key & 63) >> 3) ^ (row & 7)) << 4));
}

}  // namespace

// ==========================================================================================
// forward
// ==========================================================================================
// Warp roles (384 threads): 0 Q / K TMA producer (+ TMEM allocation), 1-4 / 5-8 the two softmax groups, 9 S = Q K^T
// issuer, 10 O = P V issuer, 11 V TMA producer (its ring frees later than the Q / K ring: one in-order producer would
// hold the next Q / K loads behind a V entry that waits for a P V product).  The key bias (padding masks) of an item is
// written by its softmax group itself.  The two MMA issuers are separate warps with the HIGHEST warp ids: measured with
// one in-order issuer as warp 1, S(i+2) waited behind P(i) and every tcgen05.mma took ~150 cycles to issue next to
// two busy softmax warps on its scheduler (55 cycles alone, tools/probes/mma_probe.cu).  A group publishes P(i) and only
// then finishes item i-2 (its previous one), so the P V product never sits on its critical path.
template <bool MASKED, bool DROP, bool TRACE>
__global__ void __launch_bounds__(TC_THREADS, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const TcDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* qk_full = bars;                          // [nslot] Q and K of the item landed
  uint64_t* qk_empty = qk_full + TC_MAX_SLOTS;       // [nslot] S = Q K^T finished reading the entry
  uint64_t* v_full = qk_empty + TC_MAX_SLOTS;        // [nv] V landed
  uint64_t* v_empty = v_full + TC_MAX_SLOTS;         // [nv] P V finished reading the entry
  uint64_t* p_full = v_empty + TC_MAX_SLOTS;         // [2] P written by softmax group g
  uint64_t* p_empty = p_full + 2;                    // [2] P V finished reading group g's P buffer
  uint64_t* bias_full = p_empty + 2;                 // [2] key bias of group g's item written
  uint64_t* bias_empty = bias_full + 2;              // [2] ... and consumed
  uint64_t* s_full = bias_full + TC_MAX_SLOTS;       // [ns] S complete in TMEM
  uint64_t* s_empty = s_full + TC_MAX_TM;            // [ns] S has been read twice
  uint64_t* o_full = s_empty + TC_MAX_TM;            // [no] O complete in TMEM
  uint64_t* o_empty = o_full + TC_MAX_TM;            // [no] O has been read
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_empty + TC_MAX_TM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ns = p.ntm, no = p.n_o;
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q);
      tma_prefetch_desc(&tmap_k);
      tma_prefetch_desc(&tmap_v);
    }
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  if (warp == 9 && lane == 0) {
    for (int i = 0; i < TC_MAX_SLOTS; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&qk_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 128);
      mbar_init(&p_empty[i], 1);
      mbar_init(&bias_full[i], 32);
      mbar_init(&bias_empty[i], 128);
    }
    for (int i = 0; i < TC_MAX_TM; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 128);
    }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_o = tmem_base + ns * p.half;   // O stages follow the S stages
  pdl_wait();
  pdl_trigger();

  const int first = blockIdx.x, stride = gridDim.x;
  const int n_local = (p.n_items > first) ? (p.n_items - first + stride - 1) / stride : 0;
  const int half = p.half;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      RingPos rq;
      ItemPos it(first, stride, p.H);
#pragma unroll 1
      for (int i = 0; i < n_local; ++i, rq.advance(p.nslot), it.advance()) {
        const int b = it.b, h = it.h;
        // the Q | K entry is free again as soon as S = Q K^T of its previous item is complete, the V entry once that
        // item's P V is: Q / K run nslot items ahead of the S products, V nv items ahead of the P V products
        const int s = rq.idx;
        tc_wait(&qk_empty[s], rq.ph ^ 1u);
        tc_stamp<TRACE>(p, i, 0);
        uint8_t* slot = smem + s * p.slot_bytes;
        mbar_arrive_expect_tx(&qk_full[s], Q_BYTES + 2 * half * ROW_BYTES);
        tma_load_2d(slot, &tmap_q, &qk_full[s], h * TC_HD, p.cu_q ? p.cu_q[b] : b * p.Tq);
        tma_load_2d(slot + p.k_off, &tmap_k, &qk_full[s], h * TC_HD, b * p.Tk);
        tma_load_2d(slot + p.k_off + half * ROW_BYTES, &tmap_k, &qk_full[s], h * TC_HD, b * p.Tk + half);
      }
    }
  } else if (warp == 11) {
    // ============================ V TMA producer ==========================
    if (lane == 0) {
      RingPos rv;
      ItemPos it(first, stride, p.H);
#pragma unroll 1
      for (int i = 0; i < n_local; ++i, rv.advance(p.nv), it.advance()) {
        const int b = it.b, h = it.h;
        const int sv = rv.idx;
        tc_wait(&v_empty[sv], rv.ph ^ 1u);
        uint8_t* vbuf = smem + p.v_base + sv * p.v_bytes;
        mbar_arrive_expect_tx(&v_full[sv], 2 * half * ROW_BYTES);
        tma_load_2d(vbuf, &tmap_v, &v_full[sv], h * TC_HD, b * p.Tk);
        tma_load_2d(vbuf + half * ROW_BYTES, &tmap_v, &v_full[sv], h * TC_HD, b * p.Tk + half);
      }
    }
  } else if (warp == 9) {
    // ============================ S = Q K^T issuer =========================
    // everything per instruction is an add on a precomputed descriptor (the address field of a shared-memory
    // descriptor is its low 14 bits in 16-byte units, and shared memory ends below 2^18 bytes); the whole warp walks the
    // loop so that this arithmetic stays warp-uniform (uniform registers), one elected lane issues
    {
      const uint32_t idesc_s = make_idesc_bf16(64, half, false, false);
      const uint64_t dq0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t dk0 = make_smem_desc_sw128(smem_u32(smem) + p.k_off, 16, 1024);
      const uint64_t slot_u = static_cast<uint64_t>(p.slot_bytes >> 4);
      const uint64_t half_u = static_cast<uint64_t>((half * ROW_BYTES) >> 4);
      RingPos rq, rt;
#pragma unroll 1
      for (int i = 0; i < n_local; ++i, rq.advance(p.nslot), rt.advance(ns)) {
        const int s = rq.idx, t = rt.idx;
        tc_wait(&qk_full[s], rq.ph);
        tc_wait(&s_empty[t], rt.ph ^ 1u);
        tc_fence_after();
        const uint64_t dq = dq0 + slot_u * s;
        const uint64_t dk = dk0 + slot_u * s;
        const uint32_t d = tmem_base + t * half;
        if (elect_one()) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
            for (int k = 0; k < TC_HD / 16; ++k)
              umma_bf16(d + (static_cast<uint32_t>(hf * 16) << 16), dq + static_cast<uint64_t>(k * 2),
                        dk + half_u * hf + static_cast<uint64_t>(k * 2), idesc_s, k > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[t]);
          umma_commit(&qk_empty[s]);
          tc_stamp<TRACE>(p, i, 1);
        }
        __syncwarp();
      }
    }
  } else if (warp == 10) {
    // ============================ O = P V issuer ===========================
    {
      const uint32_t idesc_o = make_idesc_bf16(64, TC_HD, false, true);
      const uint64_t dp0 = make_smem_desc_sw128(smem_u32(smem) + p.p_base, 16, 1024);              // K-major P buffer
      const uint64_t dv0 = make_smem_desc_sw128(smem_u32(smem) + p.v_base, 64 * ROW_BYTES, 1024);  // MN-major V tile
      const uint64_t pbuf_u = static_cast<uint64_t>(p.p_bytes >> 4), vbuf_u = static_cast<uint64_t>(p.v_bytes >> 4);
      const int nslab = p.nslab, ksteps = p.Tk16 / 16;
      RingPos rp, rv, ro;
#pragma unroll 1
      for (int i = 0; i < n_local; ++i, rp.advance(p.n_p), rv.advance(p.nv), ro.advance(no)) {
        const int g = rp.idx, sv = rv.idx, t = ro.idx;
        tc_wait(&o_empty[t], ro.ph ^ 1u);
        tc_wait(&v_full[sv], rv.ph);
        tc_wait(&p_full[g], rp.ph);
        tc_fence_after();
        const uint64_t dp = dp0 + pbuf_u * g;
        const uint64_t dv = dv0 + vbuf_u * sv;
        const uint32_t d = tmem_o + t * TC_HD;
        if (elect_one()) {
          tc_stamp<TRACE>(p, i, 2);
          for (int sl = 0; sl < nslab; ++sl) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int kk = sl * 4 + q;
              if (kk < ksteps)
                umma_bf16(d, dp + static_cast<uint64_t>(sl * (SLAB_BYTES64 >> 4) + q * 2),
                          dv + static_cast<uint64_t>(kk * ((16 * ROW_BYTES) >> 4)), idesc_o, kk > 0 ? 1u : 0u);
            }
          }
          umma_commit(&o_full[t]);
          umma_commit(&v_empty[sv]);
          umma_commit(&p_empty[i & 1]);                // by parity of the item: each softmax group is the only waiter of "its" barrier
          tc_stamp<TRACE>(p, i, 3);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================ softmax groups (warps 1-4 and 5-8) =======
    const int grp = (warp - 1) >> 2;                 // 0 / 1: items of this CTA with i % 2 == grp
    const int w4 = warp & 3;                         // TMEM lane quarter this warp may access
    const int hf = lane >> 4;                        // key half held by this thread
    const int row = w4 * 16 + (lane & 15);           // query row
    const bool stamper = (w4 == 3) && lane == 0;
    const uint32_t lane_addr = static_cast<uint32_t>(w4 * 32) << 16;
    const bool dropping = DROP && p.drop.thr != 0;   // (the trace build is the DROP variant: it must also run without dropout)
    const uint32_t dkey = dropping ? drop_key(p.drop) : 0u;
    const bool writer = hf == 0 && row < p.Tq;       // O rows live in the lower 16 lanes of every quarter (packed: checked per sample)

    // O of local item i (TMEM stage i % no) -> rows scaled by 1 / row sum -> global memory; log-sum-exp
    auto finish = [&](int i, float inv, float lse) {
      const int item = first + i * stride;
      const int b = item / p.H, h = item % p.H;
      const int t = i % no;
      tc_wait(&o_full[t], static_cast<uint32_t>(i / no) & 1u);
      tc_fence_after();
      if (stamper) tc_stamp<TRACE>(p, i, 7);
      const uint32_t t_o = tmem_o + t * TC_HD + lane_addr;
      uint32_t ro[2][32];
      tmem_ld_32x32(t_o, ro[0]);
      tmem_ld_32x32(t_o + 32, ro[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&o_empty[t]);
      long long row0 = 0;                              // packed: first row of the sample, and its own row count
      bool wr = writer;
      if (p.cu_q) {
        row0 = p.cu_q[b];
        wr = writer && row < p.cu_q[b + 1] - static_cast<int>(row0);
      }
      if (wr) {
        bf16* orow = p.o + static_cast<long long>(b) * p.o_bs + (row0 + row) * p.o_ts + h * TC_HD;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(ro[cc][g * 8 + 0]) * inv, __uint_as_float(ro[cc][g * 8 + 1]) * inv);
            u.y = pack_bf16(__uint_as_float(ro[cc][g * 8 + 2]) * inv, __uint_as_float(ro[cc][g * 8 + 3]) * inv);
            u.z = pack_bf16(__uint_as_float(ro[cc][g * 8 + 4]) * inv, __uint_as_float(ro[cc][g * 8 + 5]) * inv);
            u.w = pack_bf16(__uint_as_float(ro[cc][g * 8 + 6]) * inv, __uint_as_float(ro[cc][g * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + cc * 32 + g * 8) = u;
          }
        }
        if (p.lse) p.lse[static_cast<long long>(item) * p.Tq + row] = lse;
      }
      if (stamper) tc_stamp<TRACE>(p, i, 8);
    };

    float prev_inv = 0.f, prev_lse = 0.f;
    int prev_i = -1;
    // (one `finish` call site: the loop runs one trip past the group's last item -- code size, see tc_wait)
#pragma unroll 1
    for (int i = grp;; i += 2) {
      const bool have = i < n_local;
      float cur_inv = 0.f, cur_lse = 0.f;
      if (have) {
      const int item = first + i * stride;
      const int t = i % ns;
      const uint32_t ph_t = static_cast<uint32_t>(i / ns) & 1u;
      const uint32_t bias = smem_u32(smem + p.bias_off) + static_cast<uint32_t>(grp * 2 * half + hf * half) * 4u;   // shared-space address of this thread's key bias
      if (MASKED && p.has_bias) {
        // key bias of this item (0 / -inf per key), written by the group's 128 threads between two group barriers: the
        // first one says every thread has finished reading the previous item's bias
        const uint32_t bw = smem_u32(smem + p.bias_off) + static_cast<uint32_t>(grp * 2 * half) * 4u;
        const int e = (warp - 1 - 4 * grp) * 32 + lane;
        const int bb = item / p.H;
        bool mk[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int j = e + k * 128;
          bool masked = j >= p.Tk;
          if (!masked && p.key_tokens) masked = p.key_tokens[static_cast<long long>(bb) * p.Tk + j] == p.pad_idx;
          if (!masked && p.key_pad_mask) masked = p.key_pad_mask[static_cast<long long>(bb) * p.Tk + j] != 0;
          mk[k] = masked;
        }
        named_barrier_sync(1 + grp, 128);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (e + k * 128 < p.Tk16) sts_f1(bw + static_cast<uint32_t>(e + k * 128) * 4u, mk[k] ? -INFINITY : 0.f);
        named_barrier_sync(1 + grp, 128);
      }
      tc_wait(&s_full[t], ph_t);
      tc_fence_after();
      if (stamper) tc_stamp<TRACE>(p, i, 4);
      const uint32_t t_s = tmem_base + t * half + lane_addr;
      const int key_base = hf * half;                // first key of this thread's half
      const int pb = i % p.n_p;
      const uint32_t sP = smem_u32(smem + p.p_base + pb * p.p_bytes);
      const int bh = item;
      float mx = -INFINITY, sum = 0.f, m2;

      // score of local column c (already a float) -> masked value for the maximum
      auto masked_raw = [&](float v, int c) __attribute__((always_inline)) {
        const int key = key_base + c;
        if (MASKED && p.has_bias) v += lds_f1(bias + static_cast<uint32_t>(c) * 4u);      // 0 / -inf: the additive form keeps the scale out of the max pass
        if (key >= p.Tk || (MASKED && p.causal && key > row)) v = -INFINITY;
        return v;
      };
      // eight consecutive scores -> probabilities (row sum, dropout) -> one 16-byte unit of the P operand
      // (fast: no mask of any kind touches these keys)
      auto emit8 = [&](const uint32_t (&r)[32], int g, int c0, bool fast) __attribute__((always_inline)) {      // g: 8-column group of the block (unrolled)
        const int key0 = key_base + c0;
        float pv[8];
        if (fast) {
#pragma unroll
          for (int j = 0; j < 8; ++j) pv[j] = ex2f(fmaf(__uint_as_float(r[g * 8 + j]), p.sl2, -m2));
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a = fmaf(__uint_as_float(r[g * 8 + j]), p.sl2, -m2);
            if (MASKED && p.has_bias) a += lds_f1(bias + static_cast<uint32_t>(c0 + j) * 4u);
            if (key0 + j >= p.Tk || (MASKED && p.causal && key0 + j > row)) a = -INFINITY;
            pv[j] = ex2f(a);
          }
        }
        sum += ((pv[0] + pv[1]) + (pv[2] + pv[3])) + ((pv[4] + pv[5]) + (pv[6] + pv[7]));
        if (DROP && dropping) {                       // the row sum keeps the undropped probabilities; P V sees dropout(P)
          const uint32_t pair0 = tc_drop_pair(p, bh, row, key0);
#pragma unroll
          for (int j = 0; j < 4; ++j) drop_apply2(pv[2 * j], pv[2 * j + 1], drop_rand(dkey, pair0 + j), p.drop.thr, p.drop.scale);
        }
        uint4 u;
        u.x = pack_bf16(pv[0], pv[1]);
        u.y = pack_bf16(pv[2], pv[3]);
        u.z = pack_bf16(pv[4], pv[5]);
        u.w = pack_bf16(pv[6], pv[7]);
        sts_v4(sP + p_unit_off(row, key0), u);
      };

      // Both passes are COMPACT runtime loops: straight-line blocks of 32 columns (enough independent work per block
      // to fill the issue slots between the quarter-rate exponentials: one softmax warp per scheduler has no other
      // warp to hide latencies behind) and a tail of 8-column groups.  A first version that kept all 104 scores of a
      // thread in registers was straight-line code executed once per item -- 50 KB of SASS streaming through the
      // instruction cache.  TMEM reads are cheap (tools/probes/mma_probe.cu: 29 cycles per 32-column load + wait with
      // four warps reading), so reading S twice costs less than one pass's registers.
      const int nblk = (half + 31) >> 5;              // 32-column blocks (the last one may reach past `half`: the columns
                                                      // behind an S stage are allocated TMEM, the values are skipped)
      // ---- pass 1: row maximum
      float mxa = -INFINITY, mxb = -INFINITY;
      tmem_sweep2(t_s, nblk, [&](uint32_t (&r)[32], int c) __attribute__((always_inline)) {
        if (!MASKED && c + 32 <= half && key_base + c + 32 <= p.Tk) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            mxa = fmaxf(mxa, __uint_as_float(r[j]));
            mxb = fmaxf(mxb, __uint_as_float(r[j + 1]));
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (c + g * 8 < half) {
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                mxa = fmaxf(mxa, masked_raw(__uint_as_float(r[g * 8 + j]), c + g * 8 + j));
                mxb = fmaxf(mxb, masked_raw(__uint_as_float(r[g * 8 + j + 1]), c + g * 8 + j + 1));
              }
            }
          }
        }
      });
      mx = fmaxf(mxa, mxb);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
      if (stamper) tc_stamp<TRACE>(p, i, 5);
      m2 = (mx == -INFINITY) ? 0.f : mx * p.sl2;       // fully masked row: P = 0, O = 0, lse = -inf
      // the P V product that last read this P buffer is complete: item i - 2 (this group's previous one) with two
      // buffers, item i - 1 (the other group's) with one.  P V(i) signals p_empty[i & 1], so every barrier has ONE
      // waiting group that consumes every phase (a shared barrier would advance two phases per group item and alias).
      if (p.n_p == 2) tc_wait(&p_empty[i & 1], (static_cast<uint32_t>(i >> 1) & 1u) ^ 1u);
      else if (i > 0) tc_wait(&p_empty[(i & 1) ^ 1], static_cast<uint32_t>((i - 1) >> 1) & 1u);
      // ---- pass 2: probabilities -> bf16 P operand
      tmem_sweep2(t_s, nblk, [&](uint32_t (&r)[32], int c) __attribute__((always_inline)) {
        if (!MASKED && c + 32 <= half && key_base + c + 32 <= p.Tk) {
#pragma unroll
          for (int g = 0; g < 4; ++g) emit8(r, g, c + g * 8, true);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if (c + g * 8 < half) emit8(r, g, c + g * 8, false);
        }
      });
      tc_fence_before();
      mbar_arrive(&s_empty[t]);                       // the S stage may be overwritten
      sum += __shfl_xor_sync(0xffffffffu, sum, 16);
      fence_proxy_async();                            // P (generic-proxy stores) -> visible to the tensor core's reads
      mbar_arrive(&p_full[pb]);
      if (stamper) tc_stamp<TRACE>(p, i, 6);
      if (lane == 0) tc_stamp<TRACE>(p, i, 9 + w4);          // per-warp publication time (slots 9-12, by TMEM lane quarter)
      cur_inv = sum > 0.f ? 1.f / sum : 0.f;
      cur_lse = (sum > 0.f) ? m2 * LN2 + logf(sum) : -INFINITY;
      }
      // ---- finish the PREVIOUS item of this group while the tensor core works on this one
      if (prev_i >= 0) finish(prev_i, prev_inv, prev_lse);
      if (!have) break;
      prev_i = i;
      prev_inv = cur_inv;
      prev_lse = cur_lse;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ==========================================================================================
// backward
// ==========================================================================================
// Everything is computed TRANSPOSED (thread = key) so that the softmax statistics are per COLUMN and arrive as
// broadcast shared-memory reads (no cross-lane reduction anywhere), and the unit of work is one 128-key tile of one
// (image, head) item:
//   S^T  = K_t Q^T      [128 keys x NQ rows]   A = K tile (K-major), B = Q (K-major)            -> TMEM, NQ columns
//   dP^T = V_t dO^T     [128 x NQ]             A = V tile,            B = dO                     -> TMEM, NQ columns
//   elementwise (thread = key): P^T = exp2(S^T * scale*log2e - lse2[row]), dS^T = P^T * (dP^T - delta[row]),
//                       both as bf16 into shared memory, one 128-byte swizzle row per key ([keys][rows])
//   dV_t = P^T dO       [128 keys x 64]        A = P^T (K-major), B = dO as stored (MN-major)    -> TMEM, 64 columns
//   dK_t = dS^T Q       [128 keys x 64]        A = dS^T (K-major), B = Q as stored (MN-major)    -> TMEM, 64 columns
//   dQ  += dS K_t       [64 rows x 64]         A = dS^T read MN-major, B = K tile MN-major       -> TMEM, 64 columns,
//                                              accumulated over the key tiles of the item
// TMEM: 2 stages x (48 + 48 + 64 + 64) + 64 = 512 columns, which is why this path takes NQ <= 48 query rows (the
// caption length 48 of every BASELINE config gives T = 47); longer targets keep the mma.sync kernels.
// Warps (512 threads): 0-3 / 4-7 the two elementwise groups (units alternate), 8-11 the output group (dQ of the item,
// dV and dK tiles: TMEM -> bf16 -> per-warp swizzled staging tile -> TMA store through 3-D tensor maps that clip at the
// sample's last row), 12 TMA producer, 13 row statistics (lse * log2e and delta = rowsum(dO * O) per item, from the
// TMA-loaded dO and O tiles) + TMEM allocation, 14 issuer of the S^T / dP^T products, 15 issuer of the dV / dK / dQ
// products.
// History (cfg2 cross attention, 256 x 12 items of 47 x 197; mma.sync kernel: 171 us):
//   128 us  one output group that also computed the statistics from global memory, rows stored with per-thread 16-byte
//           stores (32 different lines per instruction: the LSU was the bottleneck, and it slowed the elementwise
//           groups' shared-memory traffic with it);
//   108 us  statistics on their own warp; MMA issue by an elected lane of a CONVERGED warp (descriptor arithmetic in
//           uniform registers instead of an ELECT / 5 x R2UR "waterfall" loop per instruction);
//    96 us  output through staging tiles + TMA stores;
//   then every added feature (trace stamps, unrolled / pipelined role bodies, more statistics warps, schedule knobs)
//   made ALL roles slower: the binary had grown to 61 KB and stall_no_inst dominated -- six roles run six different
//   loops at once and share a 32 KB L1.5 instruction cache.  This version is the same schedule on a code diet:
//   seven-instruction waits, trace stamps compiled out (template), one loop body per role.
struct TcBwdDev {
  int B, H, Tq, Tk, NQ, NT, tail16, n_items;
  int causal, has_mask;
  int kring_off, vring_off, o_off, pds_off, stage_off, stat_off, bar_off;   // byte offsets (Q | dO slots start at 0)
  const long long* key_tokens; long long pad_idx;
  const unsigned char* key_pad_mask;
  float scale, sl2;
  const float* lse;
  const int32_t* cu_q;                            // packed query rows (attention.cuh); dQ then leaves through plain row stores
  bf16* dq; long long dq_ts;
  DropCfg drop;                                   // dropout on the probabilities (functional.py:6682): same counter-based masks as the forward
  long long* trace;                               // bring-up instrument: [unit][16] clock64 stamps of CTA 0 (null = off)
};

namespace {
constexpr int BW_THREADS = 512;
constexpr int BW_TILE = 128;                          // keys per unit
constexpr int BW_TILE_BYTES = BW_TILE * ROW_BYTES;    // 16 KB: one K or V tile, one P^T or dS^T buffer
constexpr int BW_STAGE_COLS = 224;                    // S^T 0, dP^T 48, dV 96, dK 160
constexpr int BW_DQ_COL = 448;
constexpr int BW_RK = 3, BW_RV = 2;                   // K ring / V ring entries (V is dead after dP^T: its ring is short)
constexpr int BW_NSLOT = 3;                           // Q | dO slots and row-statistics slots (items in flight)
constexpr int BW_QROWS_BYTES = 48 * ROW_BYTES;        // 6 KB: the <= 48 rows of Q, dO or O of one item
constexpr int BW_SLOT_BYTES = 2 * BW_QROWS_BYTES;     // Q | dO
constexpr int BW_STAGING_BYTES = 32 * ROW_BYTES;      // 4 KB: one [32 rows][128 B] tile; two per output warp
// stamps of CTA 0's first 32 units: 0 K load issued, 1 S^T / dP^T issued, 2 st_full seen, 3 P / dS buffers free,
// 4 P / dS published, 5 output products: operands ready, 6 ... issued, 7 out_full seen, 8 tiles stored, 9 statistics done
template <bool TRACE>
__device__ __forceinline__ void bw_stamp(const TcBwdDev& p, int u, int slot) {
  if constexpr (TRACE) {
    if (blockIdx.x == 0 && u < 32) p.trace[u * 16 + slot] = clock64();
  }
}
}  // namespace

template <bool TRACE, bool DROP>
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                   const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_k_tail, const __grid_constant__ CUtensorMap tmap_v,
                   const __grid_constant__ CUtensorMap tmap_v_tail, const __grid_constant__ CUtensorMap tmap_dq,
                   const __grid_constant__ CUtensorMap tmap_dk, const __grid_constant__ CUtensorMap tmap_dv,
                   const TcBwdDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* qdo_full = bars;                 // [3] Q and dO of the item landed
  uint64_t* qdo_empty = qdo_full + 4;        // [3] every product of the item has read them
  uint64_t* k_full = qdo_empty + 4;          // [3] K tile landed
  uint64_t* k_empty = k_full + 4;            // [3] S^T and dQ of the unit have read it
  uint64_t* v_full = k_empty + 4;            // [2] V tile landed
  uint64_t* v_empty = v_full + 4;            // [2] dP^T of the unit has read it
  uint64_t* st_full = v_empty + 4;           // [2] S^T and dP^T complete in TMEM
  uint64_t* st_empty = st_full + 2;          // [2] ... and read
  uint64_t* pds_full = st_empty + 2;         // [2] P^T and dS^T written to shared memory
  uint64_t* pds_empty = pds_full + 2;        // [2] ... and read by the output products
  uint64_t* out_full = pds_empty + 2;        // [2] dV and dK tile complete in TMEM
  uint64_t* out_empty = out_full + 2;        // [2] ... and read
  uint64_t* dq_full = out_empty + 2;         // [1]
  uint64_t* dq_empty = dq_full + 1;          // [1]
  uint64_t* stat_full = dq_empty + 1;        // [3] row statistics of the item written
  uint64_t* stat_empty = stat_full + 4;      // [3] ... and read by all its units
  uint64_t* o_full = stat_empty + 4;         // [2] O of the item landed (it lives only until the statistics are done; with ONE
  uint64_t* o_empty = o_full + 2;            // [2] slot the chain stats(i) -> load O(i+1) -> stats(i+1) set the item period: 4.9 k cycles)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // shared memory starts out as zeros: K / V rows past the last key, P^T / dS^T row slots past NQ and the tail of a
  // short last tile are multiplied by zeros later, so they have to be finite
  for (int i = threadIdx.x; i < p.stat_off / 16; i += BW_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  if (warp == 13) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  if (warp == 12 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&stat_full[i], 32);
      mbar_init(&stat_empty[i], 128 * p.NT);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st_full[i], 1);
      mbar_init(&st_empty[i], 128);
      mbar_init(&pds_full[i], 128);
      mbar_init(&pds_empty[i], 1);
      mbar_init(&out_full[i], 1);
      mbar_init(&out_empty[i], 128);
    }
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 32);
    }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_trigger();

  const int first = blockIdx.x, stride = gridDim.x;
  const int n_local = (p.n_items > first) ? (p.n_items - first + stride - 1) / stride : 0;
  const int NT = p.NT;
  const int n_units = n_local * NT;
  const int qsteps = p.NQ / 16;
  uint8_t* const kring = smem + p.kring_off;
  uint8_t* const vring = smem + p.vring_off;
  uint8_t* const pds = smem + p.pds_off;
  const uint32_t stat_u = smem_u32(smem + p.stat_off);       // [3][lse2 64 | delta 64]

  if (warp == 12) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int u = 0;
#pragma unroll 1
      for (int i = 0; i < n_local; ++i) {
        const int item = first + i * stride;
        const int b = item / p.H, h = item % p.H;
        const int sl = i % BW_NSLOT;
        tc_wait(&qdo_empty[sl], (static_cast<uint32_t>(i / BW_NSLOT) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&qdo_full[sl], 2 * p.NQ * ROW_BYTES);
        const int qrow0 = p.cu_q ? p.cu_q[b] : b * p.Tq;
        tma_load_2d(smem + sl * BW_SLOT_BYTES, &tmap_q, &qdo_full[sl], h * TC_HD, qrow0);
        tma_load_2d(smem + sl * BW_SLOT_BYTES + BW_QROWS_BYTES, &tmap_do, &qdo_full[sl], h * TC_HD, qrow0);
        tc_wait(&o_empty[i & 1], (static_cast<uint32_t>(i >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&o_full[i & 1], p.NQ * ROW_BYTES);
        tma_load_2d(smem + p.o_off + (i & 1) * BW_QROWS_BYTES, &tmap_o, &o_full[i & 1], h * TC_HD, qrow0);
#pragma unroll 1
        for (int kt = 0; kt < NT; ++kt, ++u) {
          const int rk = u % BW_RK, rv = u % BW_RV;
          const bool tail = kt == NT - 1;
          const int rows = tail ? p.tail16 : BW_TILE;
          tc_wait(&k_empty[rk], (static_cast<uint32_t>(u / BW_RK) & 1u) ^ 1u);
          bw_stamp<TRACE>(p, u, 0);
          mbar_arrive_expect_tx(&k_full[rk], rows * ROW_BYTES);
          tma_load_2d(kring + rk * BW_TILE_BYTES, tail ? &tmap_k_tail : &tmap_k, &k_full[rk], h * TC_HD, b * p.Tk + kt * BW_TILE);
          tc_wait(&v_empty[rv], (static_cast<uint32_t>(u / BW_RV) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&v_full[rv], rows * ROW_BYTES);
          tma_load_2d(vring + rv * BW_TILE_BYTES, tail ? &tmap_v_tail : &tmap_v, &v_full[rv], h * TC_HD, b * p.Tk + kt * BW_TILE);
        }
      }
    }
  } else if (warp == 14) {
    // ============================ S^T = K Q^T and dP^T = V dO^T ============
    // The WHOLE warp walks the loop (waits, descriptor arithmetic: warp-uniform, so it stays in uniform registers) and
    // one elected lane issues.  Issuing from inside an `if (lane == 0)` region made every tcgen05.mma a ~19-instruction
    // sequence: 64-bit adds in vector registers plus an ELECT / 5 x R2UR "waterfall" loop per instruction.
    const uint32_t idesc = make_idesc_bf16(128, p.NQ, false, false);
    const uint64_t d_q0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t d_k0 = make_smem_desc_sw128(smem_u32(kring), 16, 1024);
    const uint64_t d_v0 = make_smem_desc_sw128(smem_u32(vring), 16, 1024);
    int i = 0, kt = 0;
#pragma unroll 1
    for (int u = 0; u < n_units; ++u) {
      const int g = u & 1, rk = u % BW_RK, rv = u % BW_RV;
      if (kt == 0) tc_wait(&qdo_full[i % BW_NSLOT], static_cast<uint32_t>(i / BW_NSLOT) & 1u);
      tc_wait(&k_full[rk], static_cast<uint32_t>(u / BW_RK) & 1u);
      tc_wait(&v_full[rv], static_cast<uint32_t>(u / BW_RV) & 1u);
      tc_wait(&st_empty[g], (static_cast<uint32_t>(u >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint64_t dq = d_q0 + static_cast<uint64_t>((i % BW_NSLOT) * (BW_SLOT_BYTES >> 4));
      const uint64_t ddo = dq + static_cast<uint64_t>(BW_QROWS_BYTES >> 4);
      const uint64_t dk = d_k0 + static_cast<uint64_t>(rk * (BW_TILE_BYTES >> 4));
      const uint64_t dv = d_v0 + static_cast<uint64_t>(rv * (BW_TILE_BYTES >> 4));
      const uint32_t t_s = tmem_base + g * BW_STAGE_COLS;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < TC_HD / 16; ++k) umma_bf16(t_s, dk + static_cast<uint64_t>(k * 2), dq + static_cast<uint64_t>(k * 2), idesc, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < TC_HD / 16; ++k) umma_bf16(t_s + 48, dv + static_cast<uint64_t>(k * 2), ddo + static_cast<uint64_t>(k * 2), idesc, k > 0 ? 1u : 0u);
        umma_commit(&st_full[g]);
        umma_commit(&v_empty[rv]);
        bw_stamp<TRACE>(p, u, 1);
      }
      __syncwarp();
      if (++kt == NT) { kt = 0; ++i; }
    }
  } else if (warp == 15) {
    // ============================ dV = P^T dO, dK = dS^T Q, dQ += dS K ======
    const uint32_t idesc_kv = make_idesc_bf16(128, TC_HD, false, true);
    const uint32_t idesc_dq = make_idesc_bf16(64, TC_HD, true, true);
    const uint64_t d_p0 = make_smem_desc_sw128(smem_u32(pds), 16, 1024);                          // K-major P^T / dS^T
    const uint64_t d_pm0 = make_smem_desc_sw128(smem_u32(pds), 64 * ROW_BYTES, 1024);             // the same bytes, MN-major
    const uint64_t d_qm0 = make_smem_desc_sw128(smem_u32(smem), 64 * ROW_BYTES, 1024);            // Q / dO as stored (MN-major)
    const uint64_t d_km0 = make_smem_desc_sw128(smem_u32(kring), 64 * ROW_BYTES, 1024);           // K tile as stored
    constexpr uint64_t KSTEP_MN = (16 * ROW_BYTES) >> 4;                                            // 16 rows of 128 B
    int i = 0, kt = 0;
#pragma unroll 1
    for (int u = 0; u < n_units; ++u) {
      const int g = u & 1, rk = u % BW_RK;
      tc_wait(&pds_full[g], static_cast<uint32_t>(u >> 1) & 1u);
      tc_wait(&out_empty[g], (static_cast<uint32_t>(u >> 1) & 1u) ^ 1u);
      tc_fence_after();
      bw_stamp<TRACE>(p, u, 5);
      const uint64_t dp = d_p0 + static_cast<uint64_t>(g * ((2 * BW_TILE_BYTES) >> 4));
      const uint64_t dds = dp + static_cast<uint64_t>(BW_TILE_BYTES >> 4);
      const uint64_t dds_m = d_pm0 + static_cast<uint64_t>(g * ((2 * BW_TILE_BYTES) >> 4) + (BW_TILE_BYTES >> 4));
      const uint64_t dq_m = d_qm0 + static_cast<uint64_t>((i % BW_NSLOT) * (BW_SLOT_BYTES >> 4));
      const uint64_t ddo_m = dq_m + static_cast<uint64_t>(BW_QROWS_BYTES >> 4);
      const uint64_t dk_m = d_km0 + static_cast<uint64_t>(rk * (BW_TILE_BYTES >> 4));
      const uint32_t t_v = tmem_base + g * BW_STAGE_COLS + 96, t_k = t_v + 64, t_q = tmem_base + BW_DQ_COL;
      const int ksteps = (kt == NT - 1 ? p.tail16 : BW_TILE) / 16;
      const bool last = kt == NT - 1;
      if (elect_one()) {
#pragma unroll 1
        for (int k = 0; k < qsteps; ++k) umma_bf16(t_v, dp + static_cast<uint64_t>(k * 2), ddo_m + KSTEP_MN * k, idesc_kv, k > 0 ? 1u : 0u);
#pragma unroll 1
        for (int k = 0; k < qsteps; ++k) umma_bf16(t_k, dds + static_cast<uint64_t>(k * 2), dq_m + KSTEP_MN * k, idesc_kv, k > 0 ? 1u : 0u);
        umma_commit(&out_full[g]);                     // the output group drains dV / dK while dQ is still being accumulated
      }
      __syncwarp();
      // the dQ accumulator of the previous item has been drained (only the item's first dQ product has to wait for that)
      if (kt == 0) {
        tc_wait(dq_empty, (static_cast<uint32_t>(i) & 1u) ^ 1u);
        tc_fence_after();
      }
      if (elect_one()) {
#pragma unroll 1
        for (int k = 0; k < ksteps; ++k) umma_bf16(t_q, dds_m + KSTEP_MN * k, dk_m + KSTEP_MN * k, idesc_dq, (kt > 0 || k > 0) ? 1u : 0u);
        umma_commit(&pds_empty[g]);
        umma_commit(&k_empty[rk]);
        bw_stamp<TRACE>(p, u, 6);
        if (last) {
          umma_commit(dq_full);
          umma_commit(&qdo_empty[i % BW_NSLOT]);
        }
      }
      __syncwarp();
      if (++kt == NT) { kt = 0; ++i; }
    }
  } else if (warp < 8) {
    // ============================ elementwise groups ========================
    const int g = warp >> 2, w4 = warp & 3;
    const int j = w4 * 32 + lane;                    // key inside the tile = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(w4 * 32) << 16;
    const uint32_t rowP = smem_u32(pds + g * 2 * BW_TILE_BYTES + j * ROW_BYTES);
    const int sw = j & 7;
    const bool dropping = DROP && p.drop.thr != 0;   // (the trace build is the DROP variant: it must also run without dropout)
    const uint32_t dkey = dropping ? drop_key(p.drop) : 0u;
    const uint32_t kpairs = static_cast<uint32_t>((p.Tk + 1) >> 1);    // dropout draws: one per pair of adjacent keys
    int prev_i = -1;
#pragma unroll 1
    for (int u = g; u < n_units; u += 2) {
      const int i = u / NT, kt = u - i * NT;
      const int key = kt * BW_TILE + j;
      bool valid = key < p.Tk;
      if (p.has_mask && valid) {
        const long long off = static_cast<long long>((first + i * stride) / p.H) * p.Tk + key;
        if (p.key_tokens) valid = p.key_tokens[off] != p.pad_idx;
        if (valid && p.key_pad_mask) valid = p.key_pad_mask[off] == 0;
      }
      if (i != prev_i) {
        tc_wait(&stat_full[i % BW_NSLOT], static_cast<uint32_t>(i / BW_NSLOT) & 1u);
        prev_i = i;
      }
      tc_wait(&st_full[g], static_cast<uint32_t>(u >> 1) & 1u);
      tc_fence_after();
      if (j == 0) bw_stamp<TRACE>(p, u, 2);
      tc_wait(&pds_empty[g], (static_cast<uint32_t>(u >> 1) & 1u) ^ 1u);
      if (j == 0) bw_stamp<TRACE>(p, u, 3);
      const uint32_t lse2 = stat_u + static_cast<uint32_t>(i % BW_NSLOT) * 512u;      // [64] lse * log2e, then [64] delta
      const uint32_t t_s = tmem_base + g * BW_STAGE_COLS + lane_addr;
      const int lim = (p.causal && valid) ? key : (valid ? 0 : 0x7fffffff);   // rows below `lim` get P = 0 (causal: key > row)
      // dropout (forward: O = dropout(P) V): the dV operand is dropout(P), dP is masked and rescaled before dS = P (dP - delta)
      const uint32_t pair0 = DROP ? static_cast<uint32_t>(first + i * stride) * static_cast<uint32_t>(p.Tq) * kpairs + (static_cast<uint32_t>(key) >> 1) : 0u;
      const uint32_t half_sh = (key & 1) ? 16u : 0u;
      // 16 rows per trip: S^T and dP^T -> P^T, dS^T -> two 16-byte units of this key's row in each buffer (ONE loop body)
#pragma unroll 1
      for (int c = 0; c < qsteps; ++c) {
        uint32_t rs[16], rd[16];
        tmem_ld_32x16(t_s + c * 16, rs);
        tmem_ld_32x16(t_s + 48 + c * 16, rd);
        tmem_ld_wait();
        uint32_t pk[8], dk[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 l = lds_f4(lse2 + static_cast<uint32_t>(c * 4 + q) * 16u), d = lds_f4(lse2 + 256u + static_cast<uint32_t>(c * 4 + q) * 16u);
          const float la[4] = {l.x, l.y, l.z, l.w}, da[4] = {d.x, d.y, d.z, d.w};
          float pv[4], ds[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int idx = q * 4 + e;
            float pr = ex2f(fmaf(__uint_as_float(rs[idx]), p.sl2, -la[e]));
            if (c * 16 + idx < lim) pr = 0.f;
            float dpe = __uint_as_float(rd[idx]);
            pv[e] = pr;
            if (DROP && dropping) {
              const uint32_t r = drop_rand(dkey, pair0 + static_cast<uint32_t>(c * 16 + idx) * kpairs);
              const bool keep = ((r >> half_sh) & 0xFFFFu) >= p.drop.thr;
              pv[e] = keep ? pr * p.drop.scale : 0.f;
              dpe = keep ? dpe * p.drop.scale : 0.f;
            }
            ds[e] = pr * (dpe - da[e]);
          }
          pk[q * 2] = pack_bf16(pv[0], pv[1]); pk[q * 2 + 1] = pack_bf16(pv[2], pv[3]);
          dk[q * 2] = pack_bf16(ds[0], ds[1]); dk[q * 2 + 1] = pack_bf16(ds[2], ds[3]);
        }
        const uint32_t u0 = static_cast<uint32_t>(((2 * c) ^ sw) << 4), u1 = static_cast<uint32_t>(((2 * c + 1) ^ sw) << 4);
        sts_v4(rowP + u0, make_uint4(pk[0], pk[1], pk[2], pk[3]));
        sts_v4(rowP + u1, make_uint4(pk[4], pk[5], pk[6], pk[7]));
        sts_v4(rowP + BW_TILE_BYTES + u0, make_uint4(dk[0], dk[1], dk[2], dk[3]));
        sts_v4(rowP + BW_TILE_BYTES + u1, make_uint4(dk[4], dk[5], dk[6], dk[7]));
      }
      tc_fence_before();
      mbar_arrive(&st_empty[g]);
      fence_proxy_async();
      mbar_arrive(&pds_full[g]);
      mbar_arrive(&stat_empty[i % BW_NSLOT]);
      if (j == 0) bw_stamp<TRACE>(p, u, 4);
    }
  } else if (warp < 12) {
    // ============================ output group ==============================
    const int w4 = warp & 3;
    const uint32_t lane_addr = static_cast<uint32_t>(w4 * 32) << 16;
    uint8_t* const stg = smem + p.stage_off + w4 * 2 * BW_STAGING_BYTES;     // this warp's two [32 rows][128 B] swizzled staging tiles
    const uint32_t stg_row = smem_u32(stg) + static_cast<uint32_t>(lane * ROW_BYTES);
    const int sw = lane & 7;
    int n_tiles = 0;                                  // staging tiles written so far: tile n uses buffer n % 2
    int i = 0, kt = 0;
#pragma unroll 1
    for (int u = 0; u < n_units; ++u) {
      const int g = u & 1;
      const int item = first + i * stride;
      const int b = item / p.H, h = item % p.H;
      tc_wait(&out_full[g], static_cast<uint32_t>(u >> 1) & 1u);
      tc_fence_after();
      if (w4 == 0 && lane == 0) bw_stamp<TRACE>(p, u, 7);
      const bool last = kt == NT - 1;
      const int key0 = kt * BW_TILE + w4 * 32;          // first key of this warp's 32 rows
      // tiles of the unit: 0 = dQ of the item (after its last unit; FIRST, the next item's dQ products wait for the
      // accumulator), 1 = dV, 2 = dK: 64 fp32 columns of this thread's TMEM lane -> one bf16 row of the staging tile
      // -> TMA store (ONE loop body)
#pragma unroll 1
      for (int t = last ? 0 : 1; t < 3; ++t) {
        if (t == 0) {
          tc_wait(dq_full, static_cast<uint32_t>(i) & 1u);
          tc_fence_after();
        }
        // M = 64 accumulator (dQ): rows 16w .. 16w+15 live in lanes 32w + 0..15 -> 16 staging rows per warp
        const bool active = t == 0 ? (w4 * 16 < p.Tq) : (key0 < p.Tk);
        if (active && t == 0 && p.cu_q != nullptr) {
          // packed rows: a 16-row TMA box would spill into the next sample, so dQ leaves through per-thread row stores
          const int r0 = p.cu_q[b], nq = p.cu_q[b + 1] - r0;
          const int row = w4 * 16 + (lane & 15);
          uint32_t ra[32];
#pragma unroll 1
          for (int hc = 0; hc < 2; ++hc) {
            tmem_ld_32x32(tmem_base + BW_DQ_COL + lane_addr + hc * 32, ra);
            tmem_ld_wait();
            if (lane < 16 && row < nq) {
              bf16* dst = p.dq + static_cast<long long>(r0 + row) * p.dq_ts + h * TC_HD + hc * 32;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 v;
                v.x = pack_bf16(__uint_as_float(ra[q * 8 + 0]) * p.scale, __uint_as_float(ra[q * 8 + 1]) * p.scale);
                v.y = pack_bf16(__uint_as_float(ra[q * 8 + 2]) * p.scale, __uint_as_float(ra[q * 8 + 3]) * p.scale);
                v.z = pack_bf16(__uint_as_float(ra[q * 8 + 4]) * p.scale, __uint_as_float(ra[q * 8 + 5]) * p.scale);
                v.w = pack_bf16(__uint_as_float(ra[q * 8 + 6]) * p.scale, __uint_as_float(ra[q * 8 + 7]) * p.scale);
                *reinterpret_cast<uint4*>(dst + q * 8) = v;
              }
            }
          }
        } else if (active) {
          const uint32_t taddr = (t == 0 ? tmem_base + BW_DQ_COL : tmem_base + g * BW_STAGE_COLS + 96 + (t - 1) * 64) + lane_addr;
          const float sc = t == 1 ? 1.f : p.scale;
          uint32_t ra[32], rb[32];
          tmem_ld_32x32(taddr, ra);
          tmem_ld_32x32(taddr + 32, rb);
          // the buffer may be rewritten once the store issued two tiles ago has READ it (one store may still be in flight)
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          tmem_ld_wait();
          if (t != 0 || lane < 16) {
            const uint32_t dst_row = stg_row + static_cast<uint32_t>(n_tiles & 1) * BW_STAGING_BYTES;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 v;
              v.x = pack_bf16(__uint_as_float(ra[q * 8 + 0]) * sc, __uint_as_float(ra[q * 8 + 1]) * sc);
              v.y = pack_bf16(__uint_as_float(ra[q * 8 + 2]) * sc, __uint_as_float(ra[q * 8 + 3]) * sc);
              v.z = pack_bf16(__uint_as_float(ra[q * 8 + 4]) * sc, __uint_as_float(ra[q * 8 + 5]) * sc);
              v.w = pack_bf16(__uint_as_float(ra[q * 8 + 6]) * sc, __uint_as_float(ra[q * 8 + 7]) * sc);
              sts_v4(dst_row + ((q ^ sw) << 4), v);
              v.x = pack_bf16(__uint_as_float(rb[q * 8 + 0]) * sc, __uint_as_float(rb[q * 8 + 1]) * sc);
              v.y = pack_bf16(__uint_as_float(rb[q * 8 + 2]) * sc, __uint_as_float(rb[q * 8 + 3]) * sc);
              v.z = pack_bf16(__uint_as_float(rb[q * 8 + 4]) * sc, __uint_as_float(rb[q * 8 + 5]) * sc);
              v.w = pack_bf16(__uint_as_float(rb[q * 8 + 6]) * sc, __uint_as_float(rb[q * 8 + 7]) * sc);
              sts_v4(dst_row + (((4 + q) ^ sw) << 4), v);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(t == 0 ? &tmap_dq : (t == 1 ? &tmap_dv : &tmap_dk), stg + (n_tiles & 1) * BW_STAGING_BYTES, h * TC_HD,
                         t == 0 ? w4 * 16 : key0, b);
            bulk_commit();
          }
          ++n_tiles;
        }
        if (t == 0) {
          tc_fence_before();
          mbar_arrive(dq_empty);
        }
      }
      tc_fence_before();
      mbar_arrive(&out_empty[g]);
      if (last) { kt = 0; ++i; } else ++kt;
      if (w4 == 0 && lane == 0) bw_stamp<TRACE>(p, u, 8);
    }
    if (lane == 0) bulk_wait_all();
    __syncwarp();
  } else if (warp == 13) {
    // ============================ row statistics ============================
    // lse * log2(e) (+inf for rows past Tq and for fully masked rows: P = 0) and delta = rowsum(dO * O): lane = row
    // (then row + 32), from the 128-byte-swizzled dO and O tiles of the item.  Compact loops on purpose (see the header).
#pragma unroll 1
    for (int i = 0; i < n_local; ++i) {
      const int item = first + i * stride;
      const int sl = i % BW_NSLOT;
      const int nq = p.cu_q ? p.cu_q[item / p.H + 1] - p.cu_q[item / p.H] : p.Tq;     // rows of this sample
      tc_wait(&qdo_full[sl], static_cast<uint32_t>(i / BW_NSLOT) & 1u);
      tc_wait(&o_full[i & 1], static_cast<uint32_t>(i >> 1) & 1u);
      tc_wait(&stat_empty[sl], (static_cast<uint32_t>(i / BW_NSLOT) & 1u) ^ 1u);
      const uint32_t s_do = smem_u32(smem + sl * BW_SLOT_BYTES + BW_QROWS_BYTES), s_o = smem_u32(smem + p.o_off + (i & 1) * BW_QROWS_BYTES);
#pragma unroll 1
      for (int rr = 0; rr < 2; ++rr) {
        const int row = lane + 32 * rr;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, l = INFINITY;
        if (row < nq) {
          const float x = p.lse[static_cast<long long>(item) * p.Tq + row];
          l = (x == -INFINITY) ? INFINITY : x * LOG2E;
#pragma unroll 2
          for (int c = 0; c < 8; ++c) {
            const uint32_t off = static_cast<uint32_t>(row * ROW_BYTES + ((c ^ (row & 7)) << 4));
            const float4 o4 = lds_f4(s_o + off), d4 = lds_f4(s_do + off);
            float2 x0 = unpack_bf16(__float_as_uint(o4.x)), y0 = unpack_bf16(__float_as_uint(d4.x));
            float2 x1 = unpack_bf16(__float_as_uint(o4.y)), y1 = unpack_bf16(__float_as_uint(d4.y));
            float2 x2 = unpack_bf16(__float_as_uint(o4.z)), y2 = unpack_bf16(__float_as_uint(d4.z));
            float2 x3 = unpack_bf16(__float_as_uint(o4.w)), y3 = unpack_bf16(__float_as_uint(d4.w));
            a0 = fmaf(x0.x, y0.x, a0); a1 = fmaf(x0.y, y0.y, a1);
            a2 = fmaf(x1.x, y1.x, a2); a3 = fmaf(x1.y, y1.y, a3);
            a0 = fmaf(x2.x, y2.x, a0); a1 = fmaf(x2.y, y2.y, a1);
            a2 = fmaf(x3.x, y3.x, a2); a3 = fmaf(x3.y, y3.y, a3);
          }
        }
        sts_f1(stat_u + static_cast<uint32_t>(sl * 128 + row) * 4u, l);
        sts_f1(stat_u + static_cast<uint32_t>(sl * 128 + 64 + row) * 4u, (a0 + a1) + (a2 + a3));
      }
      mbar_arrive(&o_empty[i & 1]);
      mbar_arrive(&stat_full[sl]);
      if (lane == 0) bw_stamp<TRACE>(p, i * NT, 9);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ==========================================================================================
// host
// ==========================================================================================
static long long* g_tc_trace = nullptr;
void attn_tc_set_trace(long long* buf) { g_tc_trace = buf; }

static bool tc_enabled() {
  static const bool on = !(getenv("B200_ATTN_TC") && atoi(getenv("B200_ATTN_TC")) == 0);
  return on;
}

static bool tc_layout_ok(const void* ptr, long long bs, long long ts, int T, bool packed = false) {
  // the TMA view is a plain 2-D [B * T rows][columns] tensor: batch stride = T rows, 16-byte aligned pitch
  // (packed rows: [total rows][columns], the batch stride is not used)
  return ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ts % 8 == 0 && (packed || bs == static_cast<long long>(T) * ts);
}

bool attn_tc_supported(const AttnArgs& a) {
  if (!tc_enabled()) return false;
  if (a.hd != TC_HD || a.Tq > 64 || a.Tk > 288 || a.Tq < 1 || a.Tk < 1) return false;
  if (a.causal && a.Tk > a.Tq + 64) return false;
  if (a.cu_k != nullptr) return false;             // packed KEYS (self attention) stay on the mma.sync kernels
  if (a.cu_q != nullptr && a.total_q < 1) return false;
  // short key ranges (the caption's self attention, Tk = Tq <= 64) stay on the mma.sync kernel: the per-item pipeline
  // overhead of this kernel exceeds the work (measured 38 us against 19 us at cfg2)
  static const int min_tk = getenv("B200_ATTN_TC_MIN_TK") ? atoi(getenv("B200_ATTN_TC_MIN_TK")) : 65;
  if (a.Tk < min_tk) return false;
  return tc_layout_ok(a.q, a.q_bs, a.q_ts, a.Tq, a.cu_q != nullptr) && tc_layout_ok(a.k, a.k_bs, a.k_ts, a.Tk) &&
         tc_layout_ok(a.v, a.v_bs, a.v_ts, a.Tk) && (reinterpret_cast<uintptr_t>(a.o) & 15) == 0 && a.o_ts % 8 == 0;
}

static void tc_fill_common(const AttnArgs& a, TcDev* d) {
  memset(d, 0, sizeof(*d));
  d->B = a.B; d->H = a.H; d->Tq = a.Tq; d->Tk = a.Tk;
  d->Tk16 = (a.Tk + 15) / 16 * 16;
  d->half = d->Tk16 / 2;
  d->nslab = (d->Tk16 + 63) / 64;
  d->n_items = a.B * a.H;
  d->causal = a.causal;
  d->has_bias = (a.key_tokens != nullptr || a.key_pad_mask != nullptr) ? 1 : 0;
  d->key_tokens = reinterpret_cast<const long long*>(a.key_tokens); d->pad_idx = a.pad_idx;
  d->key_pad_mask = a.key_pad_mask;
  d->scale = a.scale; d->sl2 = a.scale * LOG2E;
  d->o = a.o; d->o_bs = a.cu_q ? 0 : a.o_bs; d->o_ts = a.o_ts; d->lse = a.lse;
  d->cu_q = a.cu_q;
  d->drop = a.drop;
}

int attn_tc_fwd(const AttnArgs& a, cudaStream_t s) {
  TcDev d;
  tc_fill_common(a, &d);
  // shared memory: ring of Q | K entries, ring of V entries, one P buffer per softmax group (whole 64-key slabs), key bias
  d.k_off = Q_BYTES;
  d.slot_bytes = Q_BYTES + ((d.Tk16 * ROW_BYTES + 1023) / 1024) * 1024;
  d.v_bytes = ((d.Tk16 * ROW_BYTES + 1023) / 1024) * 1024;
  d.p_bytes = d.nslab * SLAB_BYTES64;
  // TMEM: ns stages of S (half columns each: both key halves share columns at lane offsets 0 / 16) + 3 stages of O
  d.n_o = 3;
  d.ntm = (512 - d.n_o * TC_HD) / d.half;
  if (d.ntm > TC_MAX_TM) d.ntm = TC_MAX_TM;
  d.stage_cols = d.half;
  const int bias_bytes = d.has_bias ? ((2 * d.Tk16 * 4 + 1023) / 1024) * 1024 : 0;
  const int cap = 227 * 1024 - (1024 /* alignment slack */ + 1024 /* barriers */ + bias_bytes);
  // preference: two P buffers, two V entries, then as many Q | K entries as fit (>= 2); 257 image tokens get one P buffer
  d.n_p = 2; d.nv = 2; d.nslot = TC_MAX_SLOTS;
  auto bytes = [&]() { return d.n_p * d.p_bytes + d.nv * d.v_bytes + d.nslot * d.slot_bytes; };
  while (d.nslot > 2 && bytes() > cap) --d.nslot;
  if (bytes() > cap) d.n_p = 1;
  while (d.nslot < TC_MAX_SLOTS && d.n_p * d.p_bytes + d.nv * d.v_bytes + (d.nslot + 1) * d.slot_bytes <= cap) ++d.nslot;
  if (d.n_p * d.p_bytes + (d.nv + 1) * d.v_bytes + d.nslot * d.slot_bytes <= cap) ++d.nv;
  // bring-up overrides (smaller rings only)
  if (const char* e = getenv("B200_ATTN_TC_NP")) { if (atoi(e) == 1) d.n_p = 1; }
  if (const char* e = getenv("B200_ATTN_TC_NSLOT")) { if (atoi(e) >= 2 && atoi(e) < d.nslot) d.nslot = atoi(e); }
  if (const char* e = getenv("B200_ATTN_TC_NS")) { if (atoi(e) >= 2 && atoi(e) < d.ntm) d.ntm = atoi(e); }
  if (const char* e = getenv("B200_ATTN_TC_NV")) { if (atoi(e) >= 2 && atoi(e) < d.nv) d.nv = atoi(e); }
  B200_REQUIRE(d.ntm >= 2 && bytes() <= cap, "attention (tcgen05): Tk = %d does not fit two pipeline stages", a.Tk);
  d.v_base = d.nslot * d.slot_bytes;
  d.p_base = d.v_base + d.nv * d.v_bytes;
  d.bias_off = d.p_base + d.n_p * d.p_bytes;
  d.bar_off = d.bias_off + bias_bytes;
  const int smem = d.bar_off + 1024 + 1024;
  B200_REQUIRE(smem <= 227 * 1024, "attention (tcgen05): %d B of shared memory", smem);

  CUtensorMap tq, tk, tv;
  const uint64_t q_rows = a.cu_q ? static_cast<uint64_t>(a.total_q) : static_cast<uint64_t>(a.B) * a.Tq;
  if (int rc = make_tmap_2d_bf16(&tq, a.q, static_cast<uint64_t>(a.H) * TC_HD, q_rows, a.q_ts * 2, TC_HD, 64)) return rc;
  if (int rc = make_tmap_2d_bf16(&tk, a.k, static_cast<uint64_t>(a.H) * TC_HD, static_cast<uint64_t>(a.B) * a.Tk, a.k_ts * 2, TC_HD, d.half)) return rc;
  if (int rc = make_tmap_2d_bf16(&tv, a.v, static_cast<uint64_t>(a.H) * TC_HD, static_cast<uint64_t>(a.B) * a.Tk, a.v_ts * 2, TC_HD, d.half)) return rc;

  const int sms = device_sm_count();
  const int grid = d.n_items < sms ? d.n_items : sms;
  d.trace = g_tc_trace;
  const bool masked = d.has_bias || d.causal, drop = d.drop.thr != 0;
  typedef void (*Kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcDev);
  static const Kern kerns[5] = {attn_tc_fwd_kernel<false, false, false>, attn_tc_fwd_kernel<true, false, false>,
                                attn_tc_fwd_kernel<false, true, false>, attn_tc_fwd_kernel<true, true, false>,
                                attn_tc_fwd_kernel<true, true, true>};     // the trace build takes the general path
  static bool configured = false;
  if (!configured) {
    for (int i = 0; i < 5; ++i) B200_CHECK_CUDA(cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  Kern kern = d.trace != nullptr ? kerns[4] : kerns[(masked ? 1 : 0) + (drop ? 2 : 0)];
  B200_CHECK_CUDA(launch_kernel(kern, dim3(grid), dim3(TC_THREADS), static_cast<size_t>(smem), s, true, 1, tq, tk, tv, d));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static bool tc_bwd_enabled() {
  static const bool on = !(getenv("B200_ATTN_TC_BWD") && atoi(getenv("B200_ATTN_TC_BWD")) == 0);
  return on;
}

bool attn_tc_bwd_supported(const AttnArgs& a, const AttnGrads& g) {
  if (!tc_enabled() || !tc_bwd_enabled()) return false;
  if (a.hd != TC_HD || a.Tq > 48 || a.Tq < 1 || a.Tk < 1 || a.cu_k != nullptr) return false;
  // short key ranges (the caption's self attention, Tk = Tq = 47: ONE 128-key tile per item) stay on the mma.sync kernel
  // by default.  Since the O tile got a second slot this kernel is the faster one in isolation (51 us against 60 us at
  // cfg2; with one slot the chain statistics(i) -> load O(i+1) -> statistics(i+1) set the item period: 66 us), but the
  // step time did not move in same-box A/Bs (8.28 vs 8.24 / 8.28 ms) and the packed (cu_k) self attention runs on the
  // mma.sync kernel anyway: one kernel family for both layouts keeps packed and padded steps bit-identical per sample.
  static const int min_tk = getenv("B200_ATTN_TC_BWD_MIN_TK") ? atoi(getenv("B200_ATTN_TC_BWD_MIN_TK")) : 65;
  if (a.Tk < min_tk) return false;
  auto out_ok = [](const void* ptr, long long ts) { return ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ts % 8 == 0; };
  const bool pk = a.cu_q != nullptr;
  return tc_layout_ok(a.q, a.q_bs, a.q_ts, a.Tq, pk) && tc_layout_ok(a.k, a.k_bs, a.k_ts, a.Tk) && tc_layout_ok(a.v, a.v_bs, a.v_ts, a.Tk) &&
         tc_layout_ok(g.d_o, g.do_bs, g.do_ts, a.Tq, pk) && tc_layout_ok(a.o, a.o_bs, a.o_ts, a.Tq, pk) && a.lse != nullptr &&
         out_ok(g.dq, g.dq_ts) && out_ok(g.dk, g.dk_ts) && out_ok(g.dv, g.dv_ts) && g.dq_bs % 8 == 0 && g.dk_bs % 8 == 0 && g.dv_bs % 8 == 0;
}

int attn_tc_bwd(const AttnArgs& a, const AttnGrads& g, cudaStream_t s) {
  TcBwdDev d;
  memset(&d, 0, sizeof(d));
  d.B = a.B; d.H = a.H; d.Tq = a.Tq; d.Tk = a.Tk;
  d.NQ = (a.Tq + 15) / 16 * 16;
  const int tk16 = (a.Tk + 15) / 16 * 16;
  d.NT = (tk16 + BW_TILE - 1) / BW_TILE;
  d.tail16 = tk16 - (d.NT - 1) * BW_TILE;
  d.n_items = a.B * a.H;
  d.causal = a.causal;
  d.has_mask = (a.key_tokens != nullptr || a.key_pad_mask != nullptr) ? 1 : 0;
  d.key_tokens = reinterpret_cast<const long long*>(a.key_tokens); d.pad_idx = a.pad_idx;
  d.key_pad_mask = a.key_pad_mask;
  d.scale = a.scale; d.sl2 = a.scale * LOG2E;
  d.lse = a.lse;
  d.cu_q = a.cu_q; d.dq = g.dq; d.dq_ts = g.dq_ts;
  d.drop = a.drop;
  // shared memory: 3 x (Q | dO), 3 K tiles, 2 V tiles, 2 x O, 2 x (P^T | dS^T), 4 x 2 staging tiles, row statistics, barriers
  d.kring_off = BW_NSLOT * BW_SLOT_BYTES;
  d.vring_off = d.kring_off + BW_RK * BW_TILE_BYTES;
  d.o_off = d.vring_off + BW_RV * BW_TILE_BYTES;
  d.pds_off = d.o_off + 2 * BW_QROWS_BYTES;
  d.stage_off = d.pds_off + 2 * 2 * BW_TILE_BYTES;
  d.stat_off = d.stage_off + 4 * 2 * BW_STAGING_BYTES;
  d.bar_off = d.stat_off + BW_NSLOT * 512;
  const int smem = d.bar_off + 512 + 1024;
  B200_REQUIRE(smem <= 227 * 1024, "attention backward (tcgen05): %d B of shared memory", smem);

  CUtensorMap tq, tdo, to, tk, tkt, tv, tvt, tdq, tdk, tdv;
  const uint64_t cols = static_cast<uint64_t>(a.H) * TC_HD;
  const uint64_t q_rows = a.cu_q ? static_cast<uint64_t>(a.total_q) : static_cast<uint64_t>(a.B) * a.Tq;
  if (int rc = make_tmap_2d_bf16(&tq, a.q, cols, q_rows, a.q_ts * 2, TC_HD, d.NQ)) return rc;
  if (int rc = make_tmap_2d_bf16(&tdo, g.d_o, cols, q_rows, g.do_ts * 2, TC_HD, d.NQ)) return rc;
  if (int rc = make_tmap_2d_bf16(&to, a.o, cols, q_rows, a.o_ts * 2, TC_HD, d.NQ)) return rc;
  if (int rc = make_tmap_2d_bf16(&tkt, a.k, cols, static_cast<uint64_t>(a.B) * a.Tk, a.k_ts * 2, TC_HD, d.tail16)) return rc;
  if (int rc = make_tmap_2d_bf16(&tvt, a.v, cols, static_cast<uint64_t>(a.B) * a.Tk, a.v_ts * 2, TC_HD, d.tail16)) return rc;
  tk = tkt; tv = tvt;
  if (d.NT > 1) {
    if (int rc = make_tmap_2d_bf16(&tk, a.k, cols, static_cast<uint64_t>(a.B) * a.Tk, a.k_ts * 2, TC_HD, BW_TILE)) return rc;
    if (int rc = make_tmap_2d_bf16(&tv, a.v, cols, static_cast<uint64_t>(a.B) * a.Tk, a.v_ts * 2, TC_HD, BW_TILE)) return rc;
  }
  if (!a.cu_q) {
    if (int rc = make_tmap_3d_bf16(&tdq, g.dq, cols, a.Tq, a.B, g.dq_ts * 2, g.dq_bs * 2, TC_HD, 16)) return rc;
  }
  if (int rc = make_tmap_3d_bf16(&tdk, g.dk, cols, a.Tk, a.B, g.dk_ts * 2, g.dk_bs * 2, TC_HD, 32)) return rc;
  if (int rc = make_tmap_3d_bf16(&tdv, g.dv, cols, a.Tk, a.B, g.dv_ts * 2, g.dv_bs * 2, TC_HD, 32)) return rc;
  if (a.cu_q) tdq = tdk;                                // packed rows: dQ leaves through plain stores, the map is not used
  const int sms = device_sm_count();
  const int grid = d.n_items < sms ? d.n_items : sms;
  d.trace = g_tc_trace;
  static bool configured = false;
  if (!configured) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const bool drop = d.drop.thr != 0;
  auto kern = d.trace != nullptr ? attn_tc_bwd_kernel<true, true> : (drop ? attn_tc_bwd_kernel<false, true> : attn_tc_bwd_kernel<false, false>);
  B200_CHECK_CUDA(launch_kernel(kern, dim3(grid), dim3(BW_THREADS), static_cast<size_t>(smem), s, true, 1, tq, tdo, to, tk, tkt, tv, tvt, tdq, tdk, tdv, d));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
