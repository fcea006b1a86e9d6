// Micro-probe (B200, sm_100a): issue rate / latency of back-to-back tcgen05.mma (kind::f16, cta_group::1) for the tile
// shapes the attention kernels could use -- M = 64 vs 128, N = 24 .. 208, A from shared memory (SS) -- and the cost of
// tcgen05.ld.  One CTA, one issuing thread; cycles via clock64.  Build + run:  tools/probes/run_mma_probe.sh
#include "../../multimodal_image_transformer_b200/csrc/common.cuh"
#include <stdio.h>
using namespace b200;

__global__ void __launch_bounds__(128, 1) probe(int M, int N, int nmma, int same_d, long long* out, int a_mn = 0, int b_mn = 0) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tptr;
  if (threadIdx.x == 32) {
    const uint32_t idesc = make_idesc_bf16(M, N, a_mn != 0, b_mn != 0);
    const uint64_t da = a_mn ? make_smem_desc_sw128(smem_u32(smem), 8192, 1024) : make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t db = b_mn ? make_smem_desc_sw128(smem_u32(smem) + 32 * 1024, 8192, 1024) : make_smem_desc_sw128(smem_u32(smem) + 32 * 1024, 16, 1024);
    const uint64_t sa = a_mn ? 128 : 2, sb = b_mn ? 128 : 2;     // per-k-step advance in 16-byte units
    // warm-up
    for (int i = 0; i < 4; ++i) umma_bf16(tbase, da, db, idesc, i > 0);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    tc_fence_after();
    const long long t0 = clock64();
    for (int i = 0; i < nmma; ++i)
      umma_bf16(tbase + (same_d ? 0 : (i & 1) * 256), da + sa * static_cast<uint64_t>(i & 3), db + sb * static_cast<uint64_t>(i & 3), idesc, (i > 1) ? 1u : 0u);
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  // tcgen05.ld cost: 4 warps x nld loads of 32 columns, back to back
  tc_fence_after();
  {
    uint32_t r[32];
    uint32_t acc = 0;
    const uint32_t ta = tbase + (static_cast<uint32_t>(warp * 32) << 16);
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < 16; ++i) {
      tmem_ld_32x32(ta + (i & 7) * 32, r);
      tmem_ld_wait();
      acc += r[0] ^ r[31];
    }
    const long long t1 = clock64();
    if (lane == 0) out[2 + warp] = t1 - t0;
    if (acc == 0x12345) out[7] = acc;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// tcgen05.ld latency while the tensor core is busy: thread 32 keeps issuing MMAs (M = 64, N = n_mma_cols, accumulate)
// into columns [256, 256 + N) while warps 0..3 time 16 x (ld 32 columns + wait) on columns [0, 256)
__global__ void __launch_bounds__(160, 1) probe_contend(int N, int busy, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  if (threadIdx.x == 128) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tptr;
  if (warp == 4) {
    if (lane == 0 && busy) {
      const uint32_t idesc = make_idesc_bf16(64, N, false, false);
      const uint64_t da = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t db = make_smem_desc_sw128(smem_u32(smem) + 32 * 1024, 16, 1024);
      long long n = 0;
      while (!stop) {
        for (int i = 0; i < 8; ++i) umma_bf16(tbase + 256, da + static_cast<uint64_t>((i & 3) * 2), db + static_cast<uint64_t>((i & 3) * 2), idesc, 1u);
        n += 8;
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      out[6] = n;
    }
  } else {
    uint32_t r[32];
    uint32_t acc = 0;
    const uint32_t ta = tbase + (static_cast<uint32_t>(warp * 32) << 16);
    for (int i = 0; i < 2000; ++i) acc += i;          // let the MMA stream start
    const long long t0 = clock64();
    for (int i = 0; i < 64; ++i) {
      tmem_ld_32x32(ta + (i & 7) * 32, r);
      tmem_ld_wait();
      acc += r[0] ^ r[31];
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
    if (acc == 0x12345) out[7] = acc;
    __threadfence_block();
    if (threadIdx.x == 0) stop = 1;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
  {
    long long* d;
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(probe_contend, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int busy = 0; busy < 2; ++busy)
      for (int N : {64, 104, 208}) {
        cudaMemset(d, 0, 64);
        probe_contend<<<1, 160, 100 * 1024>>>(N, busy, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[8];
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("tcgen05.ld x32 + wait, 4 warps, MMA stream %s (M 64 N %3d): %.1f %.1f %.1f %.1f cycles per load   (mmas issued meanwhile: %lld) %s\n",
               busy ? "RUNNING" : "idle   ", N, h[0] / 64.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0, h[6], e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  }
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  {
    struct { int M, N, a, b; } cs[] = {{128, 48, 0, 0}, {128, 64, 0, 1}, {64, 64, 1, 1}, {64, 64, 0, 1}, {64, 64, 0, 0}, {128, 64, 1, 1}};
    for (auto c : cs)
      for (int nm : {4, 8, 16}) {
        cudaMemset(d, 0, 64);
        probe<<<1, 128, 100 * 1024>>>(c.M, c.N, nm, 1, d, c.a, c.b);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[8];
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("M %3d N %3d A %s B %s n %2d: issue %6lld cyc (%5.1f / mma)  done %6lld cyc (%5.1f / mma) %s\n", c.M, c.N, c.a ? "MN" : "K ",
               c.b ? "MN" : "K ", nm, h[0], (double)h[0] / nm, h[1], (double)h[1] / nm, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  }
  const int Ms[2] = {64, 128};
  const int Ns[5] = {24, 64, 104, 208, 256};
  for (int same = 1; same >= 0; --same)
    for (int mi = 0; mi < 2; ++mi)
      for (int ni = 0; ni < 5; ++ni) {
        const int M = Ms[mi], N = Ns[ni];
        if (M == 128 && N % 16) continue;
        for (int nm : {8, 32}) {
          cudaMemset(d, 0, 64);
          probe<<<1, 128, 100 * 1024>>>(M, N, nm, same, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[8];
          cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
          printf("M %3d N %3d n %2d same_d %d: issue %6lld cyc (%5.1f / mma)  done %6lld cyc (%5.1f / mma)   ld x32+wait per warp: %lld %lld %lld %lld cyc / 16  %s\n",
                 M, N, nm, same, h[0], (double)h[0] / nm, h[1], (double)h[1] / nm, h[2], h[3], h[4], h[5], e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
      }
  return 0;
}
