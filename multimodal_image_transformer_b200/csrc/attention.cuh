// Internal (C++) interface of the fused attention kernels.
#pragma once
#include "common.cuh"
#include <string.h>

namespace b200 {

struct AttnArgs {
  const bf16* q = nullptr; long long q_bs = 0, q_ts = 0;   // (batch, position) strides, elements
  const bf16* k = nullptr; long long k_bs = 0, k_ts = 0;
  const bf16* v = nullptr; long long v_bs = 0, v_ts = 0;
  bf16* o = nullptr; long long o_bs = 0, o_ts = 0;
  float* lse = nullptr;                                    // [B,H,Tq]
  int B = 0, H = 0, Tq = 0, Tk = 0, hd = 0;
  int causal = 0;
  const int64_t* key_tokens = nullptr; long long pad_idx = 0;
  const unsigned char* key_pad_mask = nullptr;
  float scale = 1.f;
  DropCfg drop = DropCfg{nullptr, 0u, 0u, 1.f};           // dropout on the probabilities (functional.py:6682)
  // Packed (var-len) sequences: cu_q / cu_k = int32 [B + 1] row offsets of every sample in q / o / dO / dq (resp.
  // k / v / dk / dv); sample b owns rows [cu[b], cu[b+1]).  Tq / Tk are then the LARGEST per-sample lengths (they size
  // the tiles and the [B, H, Tq] log-sum-exp rows) and the batch strides of those tensors are ignored.
  const int32_t* cu_q = nullptr;
  const int32_t* cu_k = nullptr;
  int total_q = 0, total_k = 0;                             // packed: rows of the whole q-side / k-side tensors (= cu[B])
};
struct AttnGrads {
  const bf16* d_o = nullptr; long long do_bs = 0, do_ts = 0;
  bf16* dq = nullptr; long long dq_bs = 0, dq_ts = 0;
  bf16* dk = nullptr; long long dk_bs = 0, dk_ts = 0;
  bf16* dv = nullptr; long long dv_bs = 0, dv_ts = 0;
};

int attn_fwd(const AttnArgs& a, cudaStream_t s);
int attn_bwd(const AttnArgs& a, const AttnGrads& g, cudaStream_t s);

}  // namespace b200
