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

// mbar_wait of common.cuh inlines its time-out report (printf argument marshalling) at every call site; these kernels
// wait in ~20 places, so the report lives out of line and the wait itself is a handful of instructions
__device__ __noinline__ void tc_wait_timeout(uint32_t parity) {
  printf("b200: attention mbarrier timeout block %d thread %d parity %u\n", blockIdx.x, threadIdx.x, parity);
  __trap();
}
__device__ int g_tc_wait_mode = 0;      // experiment switch (B200_ATTN_TC_WAIT): 0 spin, 1 nanosleep back-off, 2 suspend-time hint
__device__ __forceinline__ bool tc_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const int mode = g_tc_wait_mode;
  const long long t0 = clock64();
  while (true) {
    if (mode == 2) { if (tc_try_wait_hint(bar, parity, 1000000u)) return; }
    else if (mbar_try_wait(bar, parity)) return;
    if (mode == 1) __nanosleep(100);
    if (clock64() - t0 > 4000000000LL) tc_wait_timeout(parity);
  }
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
__device__ __forceinline__ void tc_stamp(const TcDev& p, int i, int slot) {
  if (p.trace != nullptr && blockIdx.x == 0 && i < 32) p.trace[i * 16 + slot] = clock64();
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
template <bool MASKED, bool DROP>
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
      for (int i = 0; i < n_local; ++i) {
        const int item = first + i * stride;
        const int b = item / p.H, h = item % p.H;
        // the Q | K entry is free again as soon as S = Q K^T of its previous item is complete, the V entry once that
        // item's P V is: Q / K run nslot items ahead of the S products, V nv items ahead of the P V products
        const int s = i % p.nslot;
        tc_wait(&qk_empty[s], (static_cast<uint32_t>(i / p.nslot) & 1u) ^ 1u);
        tc_stamp(p, i, 0);
        uint8_t* slot = smem + s * p.slot_bytes;
        mbar_arrive_expect_tx(&qk_full[s], Q_BYTES + 2 * half * ROW_BYTES);
        tma_load_2d(slot, &tmap_q, &qk_full[s], h * TC_HD, b * p.Tq);
        tma_load_2d(slot + p.k_off, &tmap_k, &qk_full[s], h * TC_HD, b * p.Tk);
        tma_load_2d(slot + p.k_off + half * ROW_BYTES, &tmap_k, &qk_full[s], h * TC_HD, b * p.Tk + half);
      }
    }
  } else if (warp == 11) {
    // ============================ V TMA producer ==========================
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int item = first + i * stride;
        const int b = item / p.H, h = item % p.H;
        const int sv = i % p.nv;
        tc_wait(&v_empty[sv], (static_cast<uint32_t>(i / p.nv) & 1u) ^ 1u);
        uint8_t* vbuf = smem + p.v_base + sv * p.v_bytes;
        mbar_arrive_expect_tx(&v_full[sv], 2 * half * ROW_BYTES);
        tma_load_2d(vbuf, &tmap_v, &v_full[sv], h * TC_HD, b * p.Tk);
        tma_load_2d(vbuf + half * ROW_BYTES, &tmap_v, &v_full[sv], h * TC_HD, b * p.Tk + half);
      }
    }
  } else if (warp == 9) {
    // ============================ S = Q K^T issuer =========================
    // one thread: everything per instruction is an add on a precomputed descriptor (the address field of a shared-
    // memory descriptor is its low 14 bits in 16-byte units, and shared memory ends below 2^18 bytes)
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(64, half, false, false);
      const uint64_t dq0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t dk0 = make_smem_desc_sw128(smem_u32(smem) + p.k_off, 16, 1024);
      const uint64_t slot_u = static_cast<uint64_t>(p.slot_bytes >> 4);
      const uint64_t half_u = static_cast<uint64_t>((half * ROW_BYTES) >> 4);
      for (int i = 0; i < n_local; ++i) {
        const int s = i % p.nslot, t = i % ns;
        tc_wait(&qk_full[s], static_cast<uint32_t>(i / p.nslot) & 1u);
        tc_wait(&s_empty[t], (static_cast<uint32_t>(i / ns) & 1u) ^ 1u);
        tc_fence_after();
        const uint64_t dq = dq0 + slot_u * s;
        const uint64_t dk = dk0 + slot_u * s;
        const uint32_t d = tmem_base + t * half;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
          for (int k = 0; k < TC_HD / 16; ++k)
            umma_bf16(d + (static_cast<uint32_t>(hf * 16) << 16), dq + static_cast<uint64_t>(k * 2),
                      dk + half_u * hf + static_cast<uint64_t>(k * 2), idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
        umma_commit(&qk_empty[s]);
        tc_stamp(p, i, 1);
      }
    }
  } else if (warp == 10) {
    // ============================ O = P V issuer ===========================
    if (lane == 0) {
      const uint32_t idesc_o = make_idesc_bf16(64, TC_HD, false, true);
      const uint64_t dp0 = make_smem_desc_sw128(smem_u32(smem) + p.p_base, 16, 1024);              // K-major P buffer
      const uint64_t dv0 = make_smem_desc_sw128(smem_u32(smem) + p.v_base, 64 * ROW_BYTES, 1024);  // MN-major V tile
      const uint64_t pbuf_u = static_cast<uint64_t>(p.p_bytes >> 4), vbuf_u = static_cast<uint64_t>(p.v_bytes >> 4);
      const int nslab = p.nslab, ksteps = p.Tk16 / 16;
      for (int i = 0; i < n_local; ++i) {
        const int g = i % p.n_p, sv = i % p.nv, t = i % no;
        tc_wait(&o_empty[t], (static_cast<uint32_t>(i / no) & 1u) ^ 1u);
        tc_wait(&v_full[sv], static_cast<uint32_t>(i / p.nv) & 1u);
        tc_wait(&p_full[g], static_cast<uint32_t>(i / p.n_p) & 1u);
        tc_fence_after();
        tc_stamp(p, i, 2);
        const uint64_t dp = dp0 + pbuf_u * g;
        const uint64_t dv = dv0 + vbuf_u * sv;
        const uint32_t d = tmem_o + t * TC_HD;
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
        tc_stamp(p, i, 3);
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
    const uint32_t dkey = DROP ? drop_key(p.drop) : 0u;
    const bool writer = hf == 0 && row < p.Tq;       // O rows live in the lower 16 lanes of every quarter

    // O of local item i (TMEM stage i % no) -> rows scaled by 1 / row sum -> global memory; log-sum-exp
    auto finish = [&](int i, float inv, float lse) {
      const int item = first + i * stride;
      const int b = item / p.H, h = item % p.H;
      const int t = i % no;
      tc_wait(&o_full[t], static_cast<uint32_t>(i / no) & 1u);
      tc_fence_after();
      if (stamper) tc_stamp(p, i, 7);
      const uint32_t t_o = tmem_o + t * TC_HD + lane_addr;
      uint32_t ro[2][32];
      tmem_ld_32x32(t_o, ro[0]);
      tmem_ld_32x32(t_o + 32, ro[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&o_empty[t]);
      if (writer) {
        bf16* orow = p.o + static_cast<long long>(b) * p.o_bs + static_cast<long long>(row) * p.o_ts + h * TC_HD;
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
      if (stamper) tc_stamp(p, i, 8);
    };

    float prev_inv = 0.f, prev_lse = 0.f;
    int prev_i = -1;
    for (int i = grp; i < n_local; i += 2) {
      const int item = first + i * stride;
      const int t = i % ns;
      const uint32_t ph_t = static_cast<uint32_t>(i / ns) & 1u;
      const float* bias = reinterpret_cast<const float*>(smem + p.bias_off) + grp * 2 * half + hf * half;
      if (MASKED && p.has_bias) {
        // key bias of this item (0 / -inf per key), written by the group's 128 threads between two group barriers: the
        // first one says every thread has finished reading the previous item's bias
        float* bw = reinterpret_cast<float*>(smem + p.bias_off) + grp * 2 * half;
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
          if (e + k * 128 < p.Tk16) bw[e + k * 128] = mk[k] ? -INFINITY : 0.f;
        named_barrier_sync(1 + grp, 128);
      }
      tc_wait(&s_full[t], ph_t);
      tc_fence_after();
      if (stamper) tc_stamp(p, i, 4);
      const uint32_t t_s = tmem_base + t * half + lane_addr;
      const int key_base = hf * half;                // first key of this thread's half
      const int pb = i % p.n_p;
      uint8_t* sP = smem + p.p_base + pb * p.p_bytes;
      const int bh = item;
      float mx = -INFINITY, sum = 0.f, m2;

      // score of local column c (already a float) -> masked value for the maximum
      auto masked_raw = [&](float v, int c) __attribute__((always_inline)) {
        const int key = key_base + c;
        if (MASKED && p.has_bias) v += bias[c];      // 0 / -inf: the additive form keeps the scale out of the max pass
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
            if (MASKED && p.has_bias) a += bias[c0 + j];
            if (key0 + j >= p.Tk || (MASKED && p.causal && key0 + j > row)) a = -INFINITY;
            pv[j] = ex2f(a);
          }
        }
        sum += ((pv[0] + pv[1]) + (pv[2] + pv[3])) + ((pv[4] + pv[5]) + (pv[6] + pv[7]));
        if (DROP) {                                   // the row sum keeps the undropped probabilities; P V sees dropout(P)
          const uint32_t pair0 = tc_drop_pair(p, bh, row, key0);
#pragma unroll
          for (int j = 0; j < 4; ++j) drop_apply2(pv[2 * j], pv[2 * j + 1], drop_rand(dkey, pair0 + j), p.drop.thr, p.drop.scale);
        }
        uint4 u;
        u.x = pack_bf16(pv[0], pv[1]);
        u.y = pack_bf16(pv[2], pv[3]);
        u.z = pack_bf16(pv[4], pv[5]);
        u.w = pack_bf16(pv[6], pv[7]);
        *reinterpret_cast<uint4*>(sP + p_unit_off(row, key0)) = u;
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
      if (stamper) tc_stamp(p, i, 5);
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
      if (stamper) tc_stamp(p, i, 6);
      if (lane == 0) tc_stamp(p, i, 9 + w4);          // per-warp publication time (slots 9-12, by TMEM lane quarter)
      // ---- finish the PREVIOUS item of this group while the tensor core works on this one
      if (prev_i >= 0) finish(prev_i, prev_inv, prev_lse);
      prev_i = i;
      prev_inv = sum > 0.f ? 1.f / sum : 0.f;
      prev_lse = (sum > 0.f) ? m2 * LN2 + logf(sum) : -INFINITY;
    }
    if (prev_i >= 0) finish(prev_i, prev_inv, prev_lse);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
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

static bool tc_layout_ok(const void* ptr, long long bs, long long ts, int T) {
  // the TMA view is a plain 2-D [B * T rows][columns] tensor: batch stride = T rows, 16-byte aligned pitch
  return ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ts % 8 == 0 && bs == static_cast<long long>(T) * ts;
}

bool attn_tc_supported(const AttnArgs& a) {
  if (!tc_enabled()) return false;
  if (a.hd != TC_HD || a.Tq > 64 || a.Tk > 288 || a.Tq < 1 || a.Tk < 1) return false;
  if (a.causal && a.Tk > a.Tq + 64) return false;
  return tc_layout_ok(a.q, a.q_bs, a.q_ts, a.Tq) && tc_layout_ok(a.k, a.k_bs, a.k_ts, a.Tk) &&
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
  d->o = a.o; d->o_bs = a.o_bs; d->o_ts = a.o_ts; d->lse = a.lse;
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
  if (int rc = make_tmap_2d_bf16(&tq, a.q, static_cast<uint64_t>(a.H) * TC_HD, static_cast<uint64_t>(a.B) * a.Tq, a.q_ts * 2, TC_HD, 64)) return rc;
  if (int rc = make_tmap_2d_bf16(&tk, a.k, static_cast<uint64_t>(a.H) * TC_HD, static_cast<uint64_t>(a.B) * a.Tk, a.k_ts * 2, TC_HD, d.half)) return rc;
  if (int rc = make_tmap_2d_bf16(&tv, a.v, static_cast<uint64_t>(a.H) * TC_HD, static_cast<uint64_t>(a.B) * a.Tk, a.v_ts * 2, TC_HD, d.half)) return rc;

  const int sms = device_sm_count();
  const int grid = d.n_items < sms ? d.n_items : sms;
  d.trace = g_tc_trace;
  {
    static bool mode_set = false;
    if (!mode_set) {
      const int mode = getenv("B200_ATTN_TC_WAIT") ? atoi(getenv("B200_ATTN_TC_WAIT")) : 0;
      B200_CHECK_CUDA(cudaMemcpyToSymbol(g_tc_wait_mode, &mode, sizeof(int)));
      mode_set = true;
    }
  }
  const bool masked = d.has_bias || d.causal, drop = d.drop.thr != 0;
  typedef void (*Kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcDev);
  static const Kern kerns[4] = {attn_tc_fwd_kernel<false, false>, attn_tc_fwd_kernel<true, false>,
                                attn_tc_fwd_kernel<false, true>, attn_tc_fwd_kernel<true, true>};
  static bool configured = false;
  if (!configured) {
    for (int i = 0; i < 4; ++i) B200_CHECK_CUDA(cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  Kern kern = kerns[(masked ? 1 : 0) + (drop ? 2 : 0)];
  B200_CHECK_CUDA(launch_kernel(kern, dim3(grid), dim3(TC_THREADS), static_cast<size_t>(smem), s, true, 1, tq, tk, tv, d));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
