// Internal (C++) interface of the tcgen05 / TMEM attention kernels (attention_tc.cu).  attn_fwd / attn_bwd in
// attention.cu route to them whenever attn_tc_supported() says the shape fits (head dim 64, <= 64 query rows, plain
// [B * T, columns] layouts); everything else keeps the mma.sync kernels.  B200_ATTN_TC=0 switches the route off.
#pragma once
#include "attention.cuh"

namespace b200 {

bool attn_tc_supported(const AttnArgs& a);
bool attn_tc_bwd_supported(const AttnArgs& a, const AttnGrads& g);
int attn_tc_fwd(const AttnArgs& a, cudaStream_t s);
// bring-up instrument: device buffer [32][16] of clock64 stamps written by CTA 0 of the following launches (null = off)
void attn_tc_set_trace(long long* buf);
int attn_tc_bwd(const AttnArgs& a, const AttnGrads& g, cudaStream_t s);

}  // namespace b200
