// Internal (C++) interface of the tcgen05 GEMM family; the C ABI in capi.cu wraps these.
#pragma once
#include "common.cuh"

namespace b200 {

enum GemmEpi { EPI_STD = 0, EPI_CE_FWD = 1, EPI_CE_BWD = 2, EPI_ARGMAX = 3 };

struct GemmProblem {
  int M = 0, N = 0, K = 0;
  const bf16* A = nullptr; int64_t lda = 0; bool a_mn = false;
  const bf16* B = nullptr; int64_t ldb = 0; bool b_mn = false;
  void* D = nullptr; int64_t ldd = 0; bool d_fp32 = false; bool accumulate = false;
  const float* bias = nullptr;
  const bf16* residual = nullptr; int64_t ldr = 0;
  const bf16* relu_mask = nullptr; int64_t ldm = 0;
  int act = 0;
  DropCfg drop = DropCfg{nullptr, 0u, 0u, 1.f};   // dropout of act(A B^T + bias), applied before the residual add
  float mask_scale = 1.f;                          // factor where relu_mask passes
  // partials: split-K without atomics.  D is an fp32 [split_k][Mpad][ldd] stack (Mpad = M rounded up
  // to 128); split i stores its partial product into slab i and the consumer (layernorm_reduce_fwd)
  // sums the slabs.  Used by the skinny decode GEMMs, whose latency is the per-CTA K loop.
  bool partials = false;
  int split_k = 1;   // 0 = auto
  int block_n = 0;   // 0 = auto
  bool b_evict_last = false;  // load B (the weights) with the L2 evict-last priority: generation re-reads the same weights every
                              // position while a read-once K/V stream passes through L2 (unpaired K-major B tiles only)
  bool single_cta = false;   // never pair CTAs (cta_group::1 tiles): no cluster start-up for latency-bound skinny GEMMs
  // --- fused softmax-CE / argmax epilogues (LM head) ---
  GemmEpi epi = EPI_STD;
  const int64_t* targets = nullptr; long long ignore_index = 0;
  float* part_max = nullptr;   // [M, n_tiles]
  float* part_sum = nullptr;   // [M, n_tiles]   (CE_FWD)   /  argmax index as float bits (ARGMAX)
  float* tgt_logit = nullptr;  // [M]
  const float* row_lse = nullptr;
  const float* inv_count = nullptr;
};

// Returns 0 / negative error. n_tiles_out (optional) = number of N tiles the launch used.
int gemm_launch(const GemmProblem& p, cudaStream_t stream, int* n_tiles_out = nullptr);
int gemm_check_launch(const GemmProblem& p, cudaStream_t stream);
// Caps the persistent grid of every GEMM launched by the calling thread while the guard lives (0 = no cap):
// generation with concurrent image partitions keeps its skinny GEMMs on the SMs the cross-attention
// stream of another partition leaves free (a statically scheduled persistent grid larger than that would
// wait for the stream to finish).
// b_evict_last: every launch inside the scope loads its weights with the L2 evict-last priority (see GemmProblem).
struct GemmGridCap {
  explicit GemmGridCap(int max_ctas, bool b_evict_last = false);
  ~GemmGridCap();
  GemmGridCap(const GemmGridCap&) = delete;
  GemmGridCap& operator=(const GemmGridCap&) = delete;
 private:
  int prev_;
  bool prev_hint_;
};
// Every GEMM launched by the calling thread while the guard lives gives up `drop` TMA pipeline stages (0 = as many
// stages as fit): one stage less leaves 24-48 KB of shared memory for kernels of other streams to become resident
// next to the persistent GEMM CTAs (the bias-gradient sums of the training backward).
struct GemmStageCap {
  explicit GemmStageCap(int drop);
  ~GemmStageCap();
  GemmStageCap(const GemmStageCap&) = delete;
  GemmStageCap& operator=(const GemmStageCap&) = delete;
 private:
  int prev_;
};
int gemm_num_n_tiles(int N, int block_n);
// number of non-empty K splits a launch with this (K, split_k) uses (= slabs written in partials mode)
int gemm_effective_splits(int K, int split_k);
int device_sm_count();

}  // namespace b200
