// Micro-probe (B200): throughput of the softmax inner loop pieces with ONE or TWO warps per SM sub-partition:
// ex2.approx (MUFU), FFMA + ex2 + FADD, and the full "8 scores -> bf16 P unit in swizzled shared memory" step.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { __nv_bfloat162 t = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&t); }

template <int MODE>
__global__ void probe(float* out, long long* cyc, int iters, float s, float m) {
  __shared__ __align__(16) uint8_t buf[64 * 128 * 4];
  float x[32];
  for (int j = 0; j < 32; ++j) x[j] = (threadIdx.x * 32 + j) * 1e-3f;
  float sum = 0.f;
  const int row = threadIdx.x & 63;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = ex2f(x[j]);
    } else if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { x[j] = ex2f(fmaf(x[j], s, -m)); sum += x[j]; }
    } else {
      float pv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) pv[j] = ex2f(fmaf(x[j], s, -m));
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        sum += ((pv[g*8] + pv[g*8+1]) + (pv[g*8+2] + pv[g*8+3])) + ((pv[g*8+4] + pv[g*8+5]) + (pv[g*8+6] + pv[g*8+7]));
        uint4 u;
        u.x = pack(pv[g*8], pv[g*8+1]); u.y = pack(pv[g*8+2], pv[g*8+3]); u.z = pack(pv[g*8+4], pv[g*8+5]); u.w = pack(pv[g*8+6], pv[g*8+7]);
        const int key = (it * 32 + g * 8) & 255;
        *reinterpret_cast<uint4*>(buf + (key >> 6) * 8192 + row * 128 + ((((key & 63) >> 3) ^ (row & 7)) << 4)) = u;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] += 1e-3f;
    }
  }
  const long long t1 = clock64();
  float acc = sum;
  for (int j = 0; j < 32; ++j) acc += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + buf[threadIdx.x];
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 64);
  const int iters = 64;
  for (int threads : {128, 256, 512}) {
    for (int mode = 0; mode < 3; ++mode) {
      if (mode == 0) probe<0><<<1, threads>>>(out, cyc, iters, 0.18f, 1.f);
      if (mode == 1) probe<1><<<1, threads>>>(out, cyc, iters, 0.18f, 1.f);
      if (mode == 2) probe<2><<<1, threads>>>(out, cyc, iters, 0.18f, 1.f);
      cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("warps/SMSP %d mode %d (%s): %lld cycles for %d x 32 columns -> %.1f cycles per 32-column block per warp-slot, %.2f cycles/column\n",
             threads / 128, mode, mode == 0 ? "ex2 only" : mode == 1 ? "ffma+ex2+add" : "ffma+ex2+sum+pack+sts", h, iters, (double)h / iters, (double)h / iters / 32);
    }
  }
  return 0;
}
