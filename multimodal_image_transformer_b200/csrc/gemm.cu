// tcgen05 / TMEM / TMA GEMM for sm_100a:   D[M,N] = epi(A[M,K] . B[N,K]^T)
//
// Replaces every aten::addmm / mm behind the reference decoder (nn.Linear in decoder.py:124,
// the packed in-projections of torch/nn/functional.py:5798-5875, out_proj :6690, linear1/2 of
// torch/nn/modules/transformer.py:1197-1199) and their autograd dgrad / wgrad products.
//
// Design (one CTA per SM, persistent over a static round-robin tile schedule):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of A and B into a STAGES-deep ring of
//               128-byte-swizzled shared-memory buffers, completion on "full" mbarriers;
//   warp 1      MMA issuer: one lane issues tcgen05.mma (128 x BLOCK_N x 16 per instruction,
//               bf16 -> fp32) into one of two TMEM accumulator stages; tcgen05.commit releases
//               the smem stage ("empty") and finally publishes the accumulator ("tmem_full");
//   warp 2      allocates / frees the 2*BLOCK_N TMEM columns;
//   warps 4-7   epilogue: tcgen05.ld the accumulator (thread = row, 32 columns at a time), apply
//               bias / activation / ReLU mask / residual (or the fused softmax-CE / argmax math of
//               the LM head) and store 16-byte vectors; arrive on "tmem_empty".
// Operands may be K-major (reduction dimension contiguous) or MN-major (transposed storage):
// the latter is what dgrad (B = W as stored) and wgrad (A = dY, B = X as stored) need, so no
// operand is ever transposed in memory.
#include "gemm.cuh"
#include <math.h>
#include <stdlib.h>

namespace b200 {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
static constexpr int UMMA_K = 16;
static constexpr int NUM_THREADS = 256;      // warps 0-3: TMA producer, MMA issuer, TMEM allocator, (idle); 4-7: epilogue group 0
static constexpr int MAX_THREADS = 384;      // + warps 8-11: epilogue group 1 (storing epilogues, B200_GEMM_EPI_GROUPS)
static constexpr int ACC_STAGES = 2;

struct GemmDev {
  int M, N, K;
  int num_m_tiles, num_n_tiles, split_k, kb_per_split, num_k_blocks;
  void* D; long long ldd; int d_fp32; int accumulate;
  DropCfg drop;    // thr != 0: dropout of the (bias + activation) result, before the residual add
  float mask_scale;  // factor applied where the ReLU mask passes (1 / (1 - p) of the FFN-activation dropout)
  int part_rows;   // > 0: split-K partial slabs of this many rows each (plain stores at row split*part_rows + m)
  const float* bias;
  const bf16* residual; long long ldr;
  const bf16* relu_mask; long long ldm;
  int act;
  int aux_mode;   // 0 none, 1 residual add, 2 relu mask (tile fetched by TMA through tmap_aux)
  int b_evict_last;   // B tiles loaded with the L2 evict-last priority (CL == 1, K-major B)
  int stage_drop;     // pipeline stages given up (frees shared memory for co-resident kernels of other streams)
  const long long* targets; long long ignore_index;
  float* part_max; float* part_sum; float* tgt_logit;
  const float* row_lse; const float* inv_count;
};

static constexpr int SLAB_BYTES = BLOCK_M * 128;   // epilogue staging slab: 128 rows x 128 B (64 bf16 / 32 fp32 columns)

// Dynamic shared memory (1024-byte aligned base, at most MAX_SMEM):
//   [stages x (A tile | B tile)] [2 output slabs; 4 when a residual / mask is fused: the tile's slabs land there and are
//   rewritten in place]
//   [2 bias tiles] [barriers].  The stage count is whatever fits: 6 (5 with aux) for the 256-wide pair tile.
static constexpr int MAX_SMEM = 227 * 1024;
static constexpr int MAX_STAGES = 8;
static constexpr int BAR_BYTES = 256;

template <int BLOCK_N, int CL>
struct SmemLayout {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = (BLOCK_N / CL) * BLOCK_K * 2;     // pair mode: each CTA stages half of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  __host__ __device__ static constexpr int epi_bytes(bool aux) { return (aux ? 4 : 2) * SLAB_BYTES + 2 * BLOCK_N * 4; }
  __host__ __device__ static constexpr int stages(bool aux) {
    return (MAX_SMEM - epi_bytes(aux) - BAR_BYTES) / STAGE_BYTES > MAX_STAGES ? MAX_STAGES
                                                                              : (MAX_SMEM - epi_bytes(aux) - BAR_BYTES) / STAGE_BYTES;
  }
  __host__ __device__ static constexpr int total(bool aux) { return stages(aux) * STAGE_BYTES + epi_bytes(aux) + BAR_BYTES; }
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}

// CL = 1: one CTA per 128 x BLOCK_N tile (cta_group::1).
// CL = 2: a CTA PAIR (cluster of 2, cta_group::2) per 256 x BLOCK_N tile.  Each CTA stages its own
//         128 rows of A and HALF of the B tile, the leader's single MMA thread issues 256-row
//         instructions that read both shared memories and write 128 accumulator lanes into each
//         CTA's TMEM.  Per CTA and k-block that is 32 KB of operands instead of 48 KB for the same
//         128 x 256 x 64 MACs: the L2 -> SM operand stream (measured bound of the CL = 1 kernel,
//         ~38 B/cycle/SM) shrinks by a third and the tensor core reads B from shared memory once per pair.
template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, int CL>
__global__ void __launch_bounds__(MAX_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_aux,
                    const GemmDev p) {
  using L = SmemLayout<BLOCK_N, CL>;
  constexpr uint32_t TMEM_COLS = ACC_STAGES * BLOCK_N;
  constexpr int NSH = BLOCK_N / CL;            // B rows (N) staged by this CTA
  const bool has_aux_smem = p.aux_mode != 0;
  const int STAGES = (L::stages(has_aux_smem) - p.stage_drop >= 2) ? L::stages(has_aux_smem) - p.stage_drop : 2;

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();   // 128-byte swizzle atoms need a 1024-byte aligned base
  const int epi_offset = STAGES * L::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + epi_offset + L::epi_bytes(has_aux_smem));
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;
  uint64_t* aux_full = tmem_empty + ACC_STAGES;      // [4] residual / mask slabs landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(aux_full + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Epilogue groups of 128 threads (1 or 2, from the launch's block size).  With two, group g owns the slabs of parity g
  // and the staging / residual buffer g: per tile each group converts half of the columns, which is what brings the
  // epilogue of a K = 768 product (~5.5 us per 256 x 256 tile with one group) under the tile's MMA time (4.6 us).
  const int EG = (static_cast<int>(blockDim.x) - 128) >> 7;
  const uint32_t cta_rank = (CL > 1) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_d);
    tma_prefetch_desc(&tmap_aux);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], CL);    // pair: leader's expect_tx arrive + the peer producer's remote arrive
      mbar_init(&empty_bar[i], 1);    // one tcgen05.commit arrival (multicast to both CTAs in pair mode)
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128 * EG * CL);   // every epilogue thread of the pair arrives on the leader's
    }
    for (int i = 0; i < 4; ++i) mbar_init(&aux_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CL > 1) { tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_ptr_smem, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all();   // peer barriers must be initialised before any remote arrive
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
  pdl_wait();
  pdl_trigger();

  // work items are enumerated per cluster; both CTAs of a pair walk the same sequence
  const int num_m_groups = (p.num_m_tiles + CL - 1) / CL;
  const int total_work = num_m_groups * p.num_n_tiles * p.split_k;
  const int w_begin = blockIdx.x / CL, w_stride = gridDim.x / CL;

  if (warp == 0) {
    // ============================ TMA producer (every CTA) ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = w_begin; w < total_work; w += w_stride) {
        const int split = w % p.split_k;
        const int tile = w / p.split_k;
        const int m_blk = (tile % num_m_groups) * CL + static_cast<int>(cta_rank);   // may be a ghost tile (>= num_m_tiles)
        const int n_blk = tile / num_m_groups;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if constexpr (CL == 1) {
            mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          } else {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], CL * L::STAGE_BYTES);   // both CTAs' bytes land on the leader's barrier
            else mbar_arrive_remote(&full_bar[stage], 0);
          }
          const int n_row0 = n_blk * BLOCK_N + static_cast<int>(cta_rank) * NSH;
          auto load = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
            if constexpr (CL == 1) tma_load_2d(dst, m, &full_bar[stage], c0, c1);
            else tma_load_2d_pair(dst, m, &full_bar[stage], c0, c1);
          };
          if constexpr (!A_MN) {
            load(sa, &tmap_a, kb * BLOCK_K, m_blk * BLOCK_M);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              load(sa + j * (BLOCK_K * 128), &tmap_a, m_blk * BLOCK_M + j * 64, kb * BLOCK_K);
          }
          if constexpr (!B_MN) {
            if constexpr (CL == 1) {
              if (p.b_evict_last) tma_load_2d_hint(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_row0, l2_policy_evict_last());
              else load(sb, &tmap_b, kb * BLOCK_K, n_row0);
            } else {
              load(sb, &tmap_b, kb * BLOCK_K, n_row0);
            }
          } else {
#pragma unroll
            for (int j = 0; j < NSH / 64; ++j)
              load(sb + j * (BLOCK_K * 128), &tmap_b, n_row0 + j * 64, kb * BLOCK_K);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA only) ==============================
    // The WHOLE warp walks the loop (waits and descriptor arithmetic are warp-uniform and stay in uniform registers) and
    // one elected lane issues: inside an `if (lane == 0)` region every tcgen05.mma cost ~21 instructions (vector-register
    // descriptor math plus an ELECT / 5 x R2UR "waterfall" loop), measured in the attention kernels.
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M * CL, BLOCK_N, A_MN, B_MN);
      constexpr uint16_t PAIR_MASK = static_cast<uint16_t>((1u << CL) - 1u);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = w_begin; w < total_work; w += w_stride) {
        const int split = w % p.split_k;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
          if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: 128-byte rows, 8-row groups 1024 B apart, 16 k-elements = +32 B.
            // MN-major: 64-element (128 B) MN rows per k, 8-k groups 1024 B apart (SBO),
            //           64-wide MN groups BLOCK_K*128 B apart (LBO), 16 k = +2048 B.
            const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                     : make_smem_desc_sw128(sa + k * (UMMA_K * 2), 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                     : make_smem_desc_sw128(sb + k * (UMMA_K * 2), 16, 1024);
            if constexpr (CL == 1) umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16_pair(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees this smem stage (in both CTAs of a pair) once the MMAs have read it
          if constexpr (CL == 1) umma_commit(&empty_bar[stage]);
          else umma_commit_pair(&empty_bar[stage], PAIR_MASK);
          // accumulator complete: publish to the epilogue warps (of both CTAs)
          if (kb == kb1 - 1) {
            if constexpr (CL == 1) umma_commit(&tmem_full[acc]);
            else umma_commit_pair(&tmem_full[acc], PAIR_MASK);
          }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ============================ epilogue ================================
    // thread = accumulator row.  Per tile: the bias slice goes to shared memory once (broadcast
    // reads afterwards); the residual / ReLU-mask tile is fetched slab by slab with TMA (128 rows x
    // 128 B, double-buffered, prefetched two slabs ahead); results are staged in 128-byte-swizzled
    // shared memory and written with TMA stores (or TMA reduce-add for split-K / accumulation),
    // which also clips the M and N tails.
    const int ew = warp & 3;                      // TMEM lane quarter this warp may read
    const int grp = (warp - 4) >> 2;              // epilogue group
    const int eall = static_cast<int>(threadIdx.x) - 128;     // index among all epilogue threads
    const int et = eall - (grp << 7);             // 0..127 = accumulator row of the tile
    const bool elected = et == 0;                 // one per group
    const uint32_t bar_free = 2u + 2u * grp, bar_staged = 3u + 2u * grp;
    // Staging slabs.  Without a fused residual / mask: two, alternating.  With one: four = the whole tile; its residual /
    // mask slabs are all requested at the top of the tile (64 KB in flight per CTA while the accumulator is still being
    // computed), land in the staging slabs, and each thread rewrites its own row in place before the TMA store.  (A
    // two-slab prefetch tied to the staging pace left every slab's load latency exposed: +38 % on the FFN dgrad.)
    uint8_t* sC = smem + epi_offset;
    float* sBias = reinterpret_cast<float*>(sC + (has_aux_smem ? 4 : 2) * SLAB_BYTES);
    constexpr bool STORES = (EPI == EPI_STD || EPI == EPI_CE_BWD);
    const bool out_f32 = (EPI == EPI_STD) && p.d_fp32;
    const int slab_cols = out_f32 ? 32 : 64;
    const int nslabs = BLOCK_N / slab_cols;
    const bool has_aux = (EPI == EPI_STD) && p.aux_mode != 0;
    const uint32_t sw = static_cast<uint32_t>(et & 7);          // 128-byte swizzle phase of this row
    int acc = 0, tile_par = 0;
    uint32_t acc_phase = 0, aux_phase = 0u;      // aux_phase: one parity bit per slab barrier
    for (int w = w_begin; w < total_work; w += w_stride) {
      const int split = w % p.split_k;
      const int tile = w / p.split_k;
      const int m_blk = (tile % num_m_groups) * CL + static_cast<int>(cta_rank);
      const int n_blk = tile / num_m_groups;
      const int m0 = m_blk * BLOCK_M;
      const int row = m0 + et;
      const bool row_ok = row < p.M;
      const int n0 = n_blk * BLOCK_N;
      const bool tile_live = m0 < p.M;             // ghost tiles of an odd pair do nothing but drain TMEM

      float* bias_s = sBias + tile_par * BLOCK_N;
      for (int i = eall; i < BLOCK_N; i += 128 * EG)
        bias_s[i] = (p.bias != nullptr && split == 0 && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
      if (has_aux && elected) {
        bulk_wait_read<0>();                       // this group's stores of the previous tile have read their slabs
        if (tile_live) {
          for (int sl = grp; sl < nslabs; sl += EG) {
            if (n0 + sl * 64 < p.N) {
              mbar_arrive_expect_tx(&aux_full[sl], SLAB_BYTES);
              tma_load_2d(sC + sl * SLAB_BYTES, &tmap_aux, &aux_full[sl], n0 + sl * 64, m0);
            }
          }
        }
      }
      named_barrier_sync(1, 128 * EG);             // bias tile visible to all epilogue threads

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(ew * 32) << 16);

      // per-row state of the fused LM-head epilogues
      uint32_t dkey = 0u;
      if constexpr (EPI == EPI_STD) {
        if (p.drop.thr) dkey = drop_key(p.drop);
      }
      float run_max = -INFINITY, run_sum = 0.f, tgt_val = 0.f;
      int best_idx = 0x7fffffff;
      long long tgt = -1;
      float lse = 0.f, gscale = 0.f;
      if constexpr (EPI == EPI_CE_FWD || EPI == EPI_CE_BWD) {
        if (row_ok) tgt = p.targets[row];
      }
      if constexpr (EPI == EPI_CE_BWD) {
        if (row_ok) {
          lse = p.row_lse[row];
          gscale = (tgt != p.ignore_index) ? __ldg(p.inv_count) : 0.f;
        }
      }

#pragma unroll 1
      for (int slab = grp; slab < nslabs; slab += EG) {
        const int b = has_aux ? slab : (slab & 1);      // (slab & 1 == grp with two groups: nslabs is even)
        const int scol0 = n0 + slab * slab_cols;
        const bool slab_live = tile_live && scol0 < p.N && p.act != 99;   // warp-uniform (act 99: diagnostic, mainloop only)
        uint8_t* sCb = sC + b * SLAB_BYTES;
        const uint8_t* sRb = sCb;                 // fused residual / mask: landed in the staging slab itself
        if constexpr (STORES) {
          if (!has_aux) named_barrier_sync(bar_free, 128);   // the TMA store that last read sC[b] has drained (see below)
        }

#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half == 1 && out_f32) break;        // fp32 output: one 32-column chunk per slab
          const int c = out_f32 ? slab : slab * 2 + half;     // 32-column chunk index within the tile
          const int col0 = n0 + c * 32;
          uint32_t r[32];
          tmem_ld_32x32(t_row + c * 32, r);       // warp-collective: executed by all lanes
          tmem_ld_wait();
          if (slab + EG >= nslabs && (half == 1 || out_f32)) {
            // last TMEM read of this tile is complete: hand the accumulator stage back to the MMA thread
            tc_fence_before();
            if constexpr (CL == 1) mbar_arrive(&tmem_empty[acc]);
            else mbar_arrive_remote(&tmem_empty[acc], 0);
          }
          if (!slab_live || col0 >= p.N) continue;

          float v[32];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + g * 4);
            v[g * 4 + 0] = __uint_as_float(r[g * 4 + 0]) + b4.x;
            v[g * 4 + 1] = __uint_as_float(r[g * 4 + 1]) + b4.y;
            v[g * 4 + 2] = __uint_as_float(r[g * 4 + 2]) + b4.z;
            v[g * 4 + 3] = __uint_as_float(r[g * 4 + 3]) + b4.w;
          }

          if constexpr (EPI == EPI_STD) {
            if (p.act == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            } else if (p.act == 2) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
            }
            if (p.drop.thr) {
              const uint32_t pair0 = (static_cast<uint32_t>(row) * static_cast<uint32_t>(p.N) + static_cast<uint32_t>(col0)) >> 1;
#pragma unroll
              for (int i = 0; i < 16; ++i) drop_apply2(v[2 * i], v[2 * i + 1], drop_rand(dkey, pair0 + i), p.drop.thr, p.drop.scale);
            }
            if (has_aux) {
              if (half == 0) {                    // slab b of the aux tile has landed
                mbar_wait(&aux_full[b], (aux_phase >> b) & 1u);
                aux_phase ^= 1u << b;
              }
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint32_t u = static_cast<uint32_t>(half * 4 + g);      // 16-byte unit within the 128-byte row
                const uint4 x = *reinterpret_cast<const uint4*>(sRb + et * 128 + ((u ^ sw) << 4));
                const uint32_t x4[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 f = unpack_bf16(x4[q]);
                  if (p.aux_mode == 1) {
                    v[g * 8 + q * 2] += f.x;
                    v[g * 8 + q * 2 + 1] += f.y;
                  } else {
                    v[g * 8 + q * 2] = (f.x > 0.f) ? v[g * 8 + q * 2] * p.mask_scale : 0.f;
                    v[g * 8 + q * 2 + 1] = (f.y > 0.f) ? v[g * 8 + q * 2 + 1] * p.mask_scale : 0.f;
                  }
                }
              }
            }
            if (!out_f32) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint4 o;
                o.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
                o.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
                o.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
                o.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
                const uint32_t u = static_cast<uint32_t>(half * 4 + g);
                *reinterpret_cast<uint4*>(sCb + et * 128 + ((u ^ sw) << 4)) = o;
              }
            } else {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                *reinterpret_cast<float4*>(sCb + et * 128 + ((static_cast<uint32_t>(g) ^ sw) << 4)) =
                    make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
            }
          } else if constexpr (EPI == EPI_CE_FWD) {
            // online softmax statistics of this row over the tile's valid columns
            float cmax = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.N) cmax = fmaxf(cmax, v[i]);
            const float new_max = fmaxf(run_max, cmax);
            float sacc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.N) sacc += __expf(v[i] - new_max);
            run_sum = run_sum * __expf(run_max - new_max) + sacc;
            run_max = new_max;
            const long long rel = tgt - col0;
            if (rel >= 0 && rel < 32) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i == rel) tgt_val = v[i];
            }
          } else if constexpr (EPI == EPI_CE_BWD) {
            const long long rel = tgt - col0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float gq = __expf(v[i] - lse);
              if (i == rel) gq -= 1.f;
              v[i] = gq * gscale;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 o;
              o.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
              o.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
              o.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
              o.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
              const uint32_t u = static_cast<uint32_t>(half * 4 + g);
              *reinterpret_cast<uint4*>(sCb + et * 128 + ((u ^ sw) << 4)) = o;
            }
          } else {  // EPI_ARGMAX: first index of the maximum (torch.argmax tie rule)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (col0 + i < p.N && v[i] > run_max) {
                run_max = v[i];
                best_idx = col0 + i;
              }
            }
          }
        }

        if constexpr (STORES) {
          fence_proxy_async();                     // staging writes -> visible to the TMA engine
          named_barrier_sync(bar_staged, 128);     // whole slab staged; everyone is done reading sR[b]
          if (elected) {
            if (slab_live) {
              if (EPI == EPI_STD && p.accumulate) tma_reduce_add_2d(&tmap_d, sCb, scol0, m0);
              else tma_store_2d(&tmap_d, sCb, scol0, m0 + split * p.part_rows);
            }
            bulk_commit();
            // one group: the store issued from the OTHER buffer has finished reading it (this thread alternates buffers);
            // two groups: this group's only buffer is staged again next, so its store must have read it
            // fused residual / mask: every slab of the tile has its own buffer, drained at the top of the next tile
            if (!has_aux) {
              if (EG == 1) bulk_wait_read<1>();
              else bulk_wait_read<0>();
            }
          }
        }
      }

      if constexpr (EPI == EPI_CE_FWD) {
        if (row_ok) {
          const long long o = static_cast<long long>(row) * p.num_n_tiles + n_blk;
          p.part_max[o] = run_max;
          p.part_sum[o] = run_sum;
          if (tgt >= n0 && tgt < n0 + BLOCK_N && tgt < p.N) p.tgt_logit[row] = tgt_val;
        }
      } else if constexpr (EPI == EPI_ARGMAX) {
        if (row_ok) {
          const long long o = static_cast<long long>(row) * p.num_n_tiles + n_blk;
          p.part_max[o] = run_max;
          p.part_sum[o] = __int_as_float(best_idx);
        }
      }
      tile_par ^= 1;
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (STORES && elected) bulk_wait_read<0>();    // shared memory must outlive the stores' reads (the writes complete with the grid)
  }

  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all();   // the pair shares shared memory, TMEM and barriers until here
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CL > 1) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(sym);
  }
  return fn;
}

static int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode_fn();
  B200_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer not 16-byte aligned");
  B200_REQUIRE((outer_stride_bytes & 15) == 0, "TMA row pitch (%llu B) not a multiple of 16",
               (unsigned long long)outer_stride_bytes);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner %llu outer %llu pitch %llu box %u x %u)",
               (int)r, (unsigned long long)inner, (unsigned long long)outer,
               (unsigned long long)outer_stride_bytes, box_inner, box_outer);
  return 0;
}
int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1) {
  PFN_encodeTiled enc = get_encode_fn();
  B200_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (stride1_bytes & 15) == 0 && (stride2_bytes & 15) == 0,
               "TMA (3-D) base / strides not 16-byte aligned");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
  return 0;
}
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, inner, outer, outer_stride_bytes, box_inner, box_outer);
}
int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                     uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, inner, outer, outer_stride_bytes, box_inner, box_outer);
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

int gemm_num_n_tiles(int N, int block_n) { return (N + block_n - 1) / block_n; }
int gemm_effective_splits(int K, int split_k) {
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  if (split_k < 1) split_k = 1;
  const int per = (num_kb + split_k - 1) / split_k;
  return (num_kb + per - 1) / per;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, int CL>
static int launch_instance(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td,
                           const CUtensorMap& tx, const GemmDev& d, int grid, cudaStream_t stream) {
  auto kern = gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN, EPI, CL>;
  static bool configured = false;
  using SL = SmemLayout<BLOCK_N, CL>;
  const bool aux = d.aux_mode != 0;
  const int n_stages = (SL::stages(aux) - d.stage_drop >= 2) ? SL::stages(aux) - d.stage_drop : 2;     // as in the kernel
  const int smem = n_stages * SL::STAGE_BYTES + SL::epi_bytes(aux) + BAR_BYTES;
  if (!configured) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    configured = true;
  }
  const bool prof = gemm_profile_enabled();
  if (prof) gemm_profile_record(stream, true, 2.0 * d.M * static_cast<double>(d.N) * d.K);
  // storing epilogues run with two epilogue groups (B200_GEMM_EPI_GROUPS=1: one); the per-row reductions of the LM-head
  // statistics / argmax epilogues walk all columns of a row in one thread and keep one
  static const int epi_groups = getenv("B200_GEMM_EPI_GROUPS") ? atoi(getenv("B200_GEMM_EPI_GROUPS")) : 2;
  const int threads = ((EPI == EPI_STD || EPI == EPI_CE_BWD) && epi_groups >= 2) ? MAX_THREADS : NUM_THREADS;
  B200_CHECK_CUDA(launch_kernel(kern, dim3(grid), dim3(threads), smem, stream, !prof, CL, ta, tb, td, tx, d));
  if (prof) gemm_profile_record(stream, false, 0.0);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static thread_local int g_grid_cap = 0;
static thread_local bool g_b_evict_last = false;
static thread_local int g_stage_cap = 0;
GemmGridCap::GemmGridCap(int max_ctas, bool b_evict_last) : prev_(g_grid_cap), prev_hint_(g_b_evict_last) {
  g_grid_cap = max_ctas > 0 ? max_ctas : 0;
  g_b_evict_last = b_evict_last;
}
GemmGridCap::~GemmGridCap() { g_grid_cap = prev_; g_b_evict_last = prev_hint_; }
GemmStageCap::GemmStageCap(int drop) : prev_(g_stage_cap) { g_stage_cap = drop > 0 ? drop : 0; }
GemmStageCap::~GemmStageCap() { g_stage_cap = prev_; }

static double wave_eff(long long work, int sms) {
  long long waves = (work + sms - 1) / sms;
  return static_cast<double>(work) / static_cast<double>(waves * sms);
}

int gemm_launch(const GemmProblem& q, cudaStream_t stream, int* n_tiles_out) {
  B200_REQUIRE(q.M > 0 && q.N > 0 && q.K > 0, "gemm: empty problem %d x %d x %d", q.M, q.N, q.K);
  B200_REQUIRE(q.A && q.B, "gemm: null operand");
  B200_REQUIRE(q.N % 8 == 0, "gemm: N (%d) must be a multiple of 8", q.N);
  B200_REQUIRE(q.epi == EPI_STD || (!q.a_mn && !q.b_mn), "gemm: fused LM-head epilogues need K-major operands");
  if (q.epi == EPI_STD || q.epi == EPI_CE_BWD) {
    B200_REQUIRE(q.D != nullptr, "gemm: null output");
    B200_REQUIRE(q.ldd % (q.d_fp32 ? 4 : 8) == 0, "gemm: ldd (%lld) breaks 16-byte store alignment", (long long)q.ldd);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(q.D) & 15) == 0, "gemm: D not 16-byte aligned");
  }
  B200_REQUIRE(!q.residual || (q.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(q.residual) & 15) == 0), "gemm: residual misaligned");
  B200_REQUIRE(!q.relu_mask || (q.ldm % 8 == 0 && (reinterpret_cast<uintptr_t>(q.relu_mask) & 15) == 0), "gemm: relu_mask misaligned");
  B200_REQUIRE(!q.bias || (reinterpret_cast<uintptr_t>(q.bias) & 15) == 0, "gemm: bias misaligned");
  B200_REQUIRE(!(q.accumulate && !q.d_fp32), "gemm: accumulate needs an fp32 output");

  const int sms = device_sm_count();
  const int num_kb = (q.K + BLOCK_K - 1) / BLOCK_K;
  const int m_tiles = (q.M + BLOCK_M - 1) / BLOCK_M;

  int block_n = q.block_n;
  int split_k = q.split_k;
  const bool may_split = (q.epi == EPI_STD) && q.d_fp32 && (q.accumulate || q.partials);
  if (split_k > 1) B200_REQUIRE(may_split, "gemm: split_k > 1 needs d_fp32 && (accumulate || partials)");
  B200_REQUIRE(!q.partials || (q.epi == EPI_STD && q.d_fp32 && !q.accumulate && !q.bias && !q.residual && !q.relu_mask && q.act == 0 && q.split_k >= 1),
               "gemm: partials mode is a bare fp32 product with an explicit split_k");
  if (block_n == 0 || split_k == 0) {
    // pick (block_n, split_k) maximising full-wave efficiency; prefer the wider tile and fewer
    // splits on ties (wider tiles halve shared-memory operand traffic per flop).
    double best = -1.0;
    int bn_best = 256, sk_best = 1;
    const int bn_opts[2] = {256, 128};
    for (int bi = 0; bi < 2; ++bi) {
      const int bn = bn_opts[bi];
      if (q.block_n != 0 && bn != q.block_n) continue;
      if (q.N <= 128 && bn == 256) continue;
      const long long tiles = static_cast<long long>(m_tiles) * gemm_num_n_tiles(q.N, bn);
      const int sk_max = (q.split_k == 0 && may_split) ? 16 : 1;
      for (int sk = 1; sk <= sk_max; sk *= 2) {
        if (q.split_k > 0 && sk != q.split_k) continue;
        if (sk > 1 && num_kb / sk < 8) break;
        const double useful = static_cast<double>(q.N) / (gemm_num_n_tiles(q.N, bn) * bn);
        double e = wave_eff(tiles * sk, sms) * useful;
        if (bn == 128) e *= 0.93;   // smem-bandwidth penalty of the narrow tile
        if (sk > 1) e *= 0.98;      // atomics
        if (e > best + 1e-9) { best = e; bn_best = bn; sk_best = sk; }
      }
    }
    if (block_n == 0) block_n = bn_best;
    if (split_k == 0) split_k = sk_best;
  }
  if (split_k < 1) split_k = 1;
  B200_REQUIRE(block_n == 128 || block_n == 256, "gemm: block_n must be 128 or 256");
  if (q.epi != EPI_STD) block_n = (q.epi == EPI_ARGMAX) ? 128 : 256;

  GemmDev d;
  d.M = q.M; d.N = q.N; d.K = q.K;
  d.num_m_tiles = m_tiles;
  d.num_n_tiles = gemm_num_n_tiles(q.N, block_n);
  d.num_k_blocks = num_kb;
  d.kb_per_split = (num_kb + split_k - 1) / split_k;
  d.split_k = (num_kb + d.kb_per_split - 1) / d.kb_per_split;  // no empty splits
  d.D = q.D; d.ldd = q.ldd; d.d_fp32 = q.d_fp32; d.accumulate = q.accumulate;
  d.part_rows = q.partials ? m_tiles * BLOCK_M : 0;
  d.drop = q.drop;
  d.mask_scale = q.mask_scale;
  B200_REQUIRE(q.drop.thr == 0 || (q.epi == EPI_STD && !q.d_fp32 && static_cast<long long>(q.M) * q.N < (1ll << 32)),
               "gemm: dropout needs the standard bf16 epilogue and M*N < 2^32");
  d.bias = q.bias; d.residual = q.residual; d.ldr = q.ldr; d.relu_mask = q.relu_mask; d.ldm = q.ldm;
  d.act = q.act;
  d.b_evict_last = (q.b_evict_last || g_b_evict_last) ? 1 : 0;
  d.stage_drop = g_stage_cap;
  d.targets = reinterpret_cast<const long long*>(q.targets); d.ignore_index = q.ignore_index;
  d.part_max = q.part_max; d.part_sum = q.part_sum; d.tgt_logit = q.tgt_logit;
  d.row_lse = q.row_lse; d.inv_count = q.inv_count;
  if (n_tiles_out) *n_tiles_out = d.num_n_tiles;

  // cluster of two CTAs sharing the B tile through TMA multicast whenever there are >= 2 M tiles
  static const bool no_cluster = getenv("B200_GEMM_NO_CLUSTER") != nullptr;
  const int cl = (!no_cluster && !q.single_cta && d.num_m_tiles >= 2 && sms >= 2) ? 2 : 1;

  CUtensorMap ta, tb;
  int rc;
  if (!q.a_mn) rc = make_tmap_2d_bf16(&ta, q.A, q.K, q.M, q.lda * 2, BLOCK_K, BLOCK_M);
  else         rc = make_tmap_2d_bf16(&ta, q.A, q.M, q.K, q.lda * 2, 64, BLOCK_K);
  if (rc) return rc;
  if (!q.b_mn) rc = make_tmap_2d_bf16(&tb, q.B, q.K, q.N, q.ldb * 2, BLOCK_K, block_n / cl);
  else         rc = make_tmap_2d_bf16(&tb, q.B, q.N, q.K, q.ldb * 2, 64, BLOCK_K);
  if (rc) return rc;
  // output and residual / mask tiles travel through 128-row x 128-byte TMA boxes
  CUtensorMap td = ta, tx = ta;
  if (q.epi == EPI_STD || q.epi == EPI_CE_BWD) {
    if (q.partials) rc = make_tmap_2d_f32(&td, q.D, q.N, static_cast<uint64_t>(d.split_k) * d.part_rows, q.ldd * 4, 32, BLOCK_M);
    else if (q.d_fp32) rc = make_tmap_2d_f32(&td, q.D, q.N, q.M, q.ldd * 4, 32, BLOCK_M);
    else          rc = make_tmap_2d_bf16(&td, q.D, q.N, q.M, q.ldd * 2, 64, BLOCK_M);
    if (rc) return rc;
  }
  d.aux_mode = 0;
  if (q.epi == EPI_STD && (q.residual || q.relu_mask)) {
    B200_REQUIRE(!(q.residual && q.relu_mask), "gemm: residual and relu_mask cannot be combined in one launch");
    B200_REQUIRE(!q.d_fp32, "gemm: residual / relu_mask need a bf16 output");
    const bf16* aux = q.residual ? q.residual : q.relu_mask;
    const long long lda_x = q.residual ? q.ldr : q.ldm;
    rc = make_tmap_2d_bf16(&tx, aux, q.N, q.M, lda_x * 2, 64, BLOCK_M);
    if (rc) return rc;
    d.aux_mode = q.residual ? 1 : 2;
  }

  const long long work = static_cast<long long>((d.num_m_tiles + cl - 1) / cl) * d.num_n_tiles * d.split_k;
  long long slots = sms / cl;
  static const int env_cap = getenv("B200_GEMM_GRID_CAP") ? atoi(getenv("B200_GEMM_GRID_CAP")) : 0;    // experiments (tools/gemm_variants.py)
  const int cap = g_grid_cap > 0 ? g_grid_cap : env_cap;
  if (cap > 0 && cap / cl >= 1 && cap / cl < slots) slots = cap / cl;
  const int grid = static_cast<int>(work < slots ? work : slots) * cl;

#define B200_GEMM_CASE(BN, AMN, BMN, EP)                                                  \
  do {                                                                                    \
    if (cl == 2) return launch_instance<BN, AMN, BMN, EP, 2>(ta, tb, td, tx, d, grid, stream);    \
    return launch_instance<BN, AMN, BMN, EP, 1>(ta, tb, td, tx, d, grid, stream);                 \
  } while (0)
  if (q.epi == EPI_CE_FWD) B200_GEMM_CASE(256, false, false, EPI_CE_FWD);
  if (q.epi == EPI_CE_BWD) B200_GEMM_CASE(256, false, false, EPI_CE_BWD);
  if (q.epi == EPI_ARGMAX) B200_GEMM_CASE(128, false, false, EPI_ARGMAX);
  if (block_n == 256) {
    if (!q.a_mn && !q.b_mn) B200_GEMM_CASE(256, false, false, EPI_STD);
    if (!q.a_mn && q.b_mn) B200_GEMM_CASE(256, false, true, EPI_STD);
    if (q.a_mn && q.b_mn) B200_GEMM_CASE(256, true, true, EPI_STD);
    if (q.a_mn && !q.b_mn) B200_GEMM_CASE(256, true, false, EPI_STD);
  } else {
    if (!q.a_mn && !q.b_mn) B200_GEMM_CASE(128, false, false, EPI_STD);
    if (!q.a_mn && q.b_mn) B200_GEMM_CASE(128, false, true, EPI_STD);
    if (q.a_mn && q.b_mn) B200_GEMM_CASE(128, true, true, EPI_STD);
    if (q.a_mn && !q.b_mn) B200_GEMM_CASE(128, true, false, EPI_STD);
  }
#undef B200_GEMM_CASE
  B200_REQUIRE(false, "gemm: unreachable dispatch");
}

// ------------------------------------------------------------------------------------------
// CUDA-core check kernel (test instrument only): same contract, fp32 accumulation of the
// bf16 operands, one thread per output element.
// ------------------------------------------------------------------------------------------
__global__ void gemm_check_kernel(GemmDev p, const bf16* A, long long lda, int a_mn, const bf16* B,
                                  long long ldb, int b_mn) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(p.M) * p.N) return;
  const int m = static_cast<int>(idx / p.N), n = static_cast<int>(idx % p.N);
  float acc = 0.f;
  for (int k = 0; k < p.K; ++k) {
    const float a = __bfloat162float(a_mn ? A[static_cast<long long>(k) * lda + m] : A[static_cast<long long>(m) * lda + k]);
    const float b = __bfloat162float(b_mn ? B[static_cast<long long>(k) * ldb + n] : B[static_cast<long long>(n) * ldb + k]);
    acc = fmaf(a, b, acc);
  }
  if (p.bias) acc += p.bias[n];
  if (p.act == 1) acc = fmaxf(acc, 0.f);
  else if (p.act == 2) acc = gelu_erf(acc);
  if (p.relu_mask) acc = (__bfloat162float(p.relu_mask[static_cast<long long>(m) * p.ldm + n]) > 0.f) ? acc * p.mask_scale : 0.f;
  if (p.residual) acc += __bfloat162float(p.residual[static_cast<long long>(m) * p.ldr + n]);
  if (p.d_fp32) {
    float* d = reinterpret_cast<float*>(p.D) + static_cast<long long>(m) * p.ldd + n;
    *d = p.accumulate ? (*d + acc) : acc;
  } else {
    reinterpret_cast<bf16*>(p.D)[static_cast<long long>(m) * p.ldd + n] = __float2bfloat16(acc);
  }
}

int gemm_check_launch(const GemmProblem& q, cudaStream_t stream) {
  B200_REQUIRE(q.epi == EPI_STD, "gemm_check: standard epilogue only");
  B200_REQUIRE(!q.partials, "gemm_check: partials mode not modelled");
  GemmDev d = {};
  d.M = q.M; d.N = q.N; d.K = q.K;
  d.D = q.D; d.ldd = q.ldd; d.d_fp32 = q.d_fp32; d.accumulate = q.accumulate;
  d.bias = q.bias; d.residual = q.residual; d.ldr = q.ldr; d.relu_mask = q.relu_mask; d.ldm = q.ldm;
  d.act = q.act;
  d.mask_scale = q.mask_scale;
  B200_REQUIRE(q.drop.thr == 0, "gemm_check: dropout not modelled");
  const long long total = static_cast<long long>(q.M) * q.N;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  gemm_check_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(d, q.A, q.lda, q.a_mn, q.B, q.ldb, q.b_mn);  // test instrument: plain launch
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
