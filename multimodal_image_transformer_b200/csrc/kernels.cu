// Memory-bound kernels of the caption decoder: every one is a single pass over its operands with
// 128-bit global accesses; reductions use warp shuffles (one warp per row for LayerNorm).
#include "kernels.cuh"
#include <math.h>
#include <stdlib.h>

namespace b200 {

static inline int cdiv(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
  o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
  return o;
}


// ------------------------------------------------------------------------------------------
// token embedding * sqrt(E) + sinusoidal PE   (decoder.py:168-170, 71-72)
// ------------------------------------------------------------------------------------------
__global__ void embed_pe_fwd_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ emb,
                                    const float* __restrict__ pe, bf16* __restrict__ x, int rows,
                                    int T, int E, int V, float scale, int t0, const DropCfg dc,
                                    const int32_t* __restrict__ pos) {
  pdl_wait();
  pdl_trigger();
  const int vec_per_row = E >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * vec_per_row) return;
  const int row = static_cast<int>(idx / vec_per_row);
  const int c = static_cast<int>(idx % vec_per_row) << 3;
  long long tok = tokens[row];
  tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
  const int t = pos ? pos[row] : row % T + t0;
  const float4* e4 = reinterpret_cast<const float4*>(emb + tok * E + c);
  const float4* p4 = reinterpret_cast<const float4*>(pe + static_cast<long long>(t) * E + c);
  const float4 e0 = __ldg(e4), e1 = __ldg(e4 + 1), p0 = __ldg(p4), p1 = __ldg(p4 + 1);
  float f[8] = {fmaf(e0.x, scale, p0.x), fmaf(e0.y, scale, p0.y), fmaf(e0.z, scale, p0.z), fmaf(e0.w, scale, p0.w),
                fmaf(e1.x, scale, p1.x), fmaf(e1.y, scale, p1.y), fmaf(e1.z, scale, p1.z), fmaf(e1.w, scale, p1.w)};
  if (dc.thr) {                                    // decoder.py:72
    const uint32_t key = drop_key(dc);
    const uint32_t pair0 = (static_cast<uint32_t>(row) * E + c) >> 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) drop_apply2(f[2 * j], f[2 * j + 1], drop_rand(key, pair0 + j), dc.thr, dc.scale);
  }
  *reinterpret_cast<uint4*>(x + static_cast<long long>(row) * E + c) = pack8(f);
}

int embed_pe_fwd(const int64_t* tokens, const float* emb, const float* pe, bf16* x, int B, int T,
                 int E, int V, float scale, cudaStream_t s, int t0, DropCfg dc, const int32_t* pos) {
  B200_REQUIRE(E % 8 == 0, "embed: E (%d) must be a multiple of 8", E);
  const long long n = static_cast<long long>(B) * T * (E / 8);
  B200_CHECK_CUDA(launch_kernel(embed_pe_fwd_kernel, dim3(cdiv(n, 256)), dim3(256), 0, s, true, 1, tokens, emb, pe, x, B * T, T, E, V, scale, t0, dc, pos));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void pack_rows_kernel(const int64_t* __restrict__ tokens, const int64_t* __restrict__ targets,
                                 const int32_t* __restrict__ cu, int B, int T, int64_t* __restrict__ ptok,
                                 int64_t* __restrict__ ptgt, int32_t* __restrict__ ppos) {
  pdl_wait();
  pdl_trigger();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * T) return;
  const int b = idx / T, t = idx - b * T;
  const int r0 = cu[b];
  if (t >= cu[b + 1] - r0) return;
  ptok[r0 + t] = tokens[idx];
  if (targets) ptgt[r0 + t] = targets[idx];
  ppos[r0 + t] = t;
}
int pack_rows(const int64_t* tokens, const int64_t* targets, const int32_t* cu, int B, int T, int64_t* ptok,
              int64_t* ptgt, int32_t* ppos, cudaStream_t s) {
  B200_CHECK_CUDA(launch_kernel(pack_rows_kernel, dim3(cdiv(static_cast<long long>(B) * T, 256)), dim3(256), 0, s, true, 1, tokens,
                                targets, cu, B, T, ptok, ptgt, ppos));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// scatter-add into the embedding gradient; the padding row receives nothing (decoder.py:105)
__global__ void embed_bwd_kernel(const int64_t* __restrict__ tokens, const bf16* __restrict__ dx,
                                 float* __restrict__ demb, int rows, int E, int V, long long pad_idx,
                                 float scale, const DropCfg dc) {
  pdl_wait();
  pdl_trigger();
  const int vec_per_row = E >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * vec_per_row) return;
  const int row = static_cast<int>(idx / vec_per_row);
  const int c = static_cast<int>(idx % vec_per_row) << 3;
  const long long tok = tokens[row];
  if (tok == pad_idx || tok < 0 || tok >= V) return;
  const uint4 g = ldg_nc_v4(dx + static_cast<long long>(row) * E + c);
  float f[8];
  unpack8(g, f);
  if (dc.thr) {
    const uint32_t key = drop_key(dc);
    const uint32_t pair0 = (static_cast<uint32_t>(row) * E + c) >> 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) drop_apply2(f[2 * j], f[2 * j + 1], drop_rand(key, pair0 + j), dc.thr, dc.scale);
  }
  float* dst = demb + tok * E + c;
  red_add_v4_f32(dst, f[0] * scale, f[1] * scale, f[2] * scale, f[3] * scale);
  red_add_v4_f32(dst + 4, f[4] * scale, f[5] * scale, f[6] * scale, f[7] * scale);
}

int embed_bwd(const int64_t* tokens, const bf16* dx, float* demb, int B, int T, int E, int V,
              long long pad_idx, float scale, cudaStream_t s, DropCfg dc) {
  B200_REQUIRE(E % 8 == 0, "embed_bwd: E (%d) must be a multiple of 8", E);
  const long long n = static_cast<long long>(B) * T * (E / 8);
  B200_CHECK_CUDA(launch_kernel(embed_bwd_kernel, dim3(cdiv(n, 256)), dim3(256), 0, s, true, 1, tokens, dx, demb, B * T, E, V, pad_idx, scale, dc));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// one draw of fresh dropout masks per training forward (state = [seed, counter], see common.cuh)
__global__ void drop_advance_kernel(uint32_t* state) {
  pdl_wait();
  pdl_trigger();
  state[1] += 1u;
}
int drop_advance(uint32_t* state, cudaStream_t s) {
  B200_CHECK_CUDA(launch_kernel(drop_advance_kernel, dim3(1), dim3(1), 0, s, true, 1, state));
  note_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------
// LayerNorm (torch.nn.LayerNorm: biased variance, eps inside the sqrt), one warp per row
// ------------------------------------------------------------------------------------------
static constexpr int LN_MAXV = 8;  // 8 x (8 bf16) x 32 lanes = E up to 2048

template <int NV>
__global__ void __launch_bounds__(128)
layernorm_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, bf16* __restrict__ y, float* __restrict__ mean,
                     float* __restrict__ rstd, int rows, int E, float eps) {
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= rows) return;
  const int nvec = E >> 3;
  const bf16* xr = x + static_cast<long long>(row) * E;
  float v[NV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      unpack8(ldg_nc_v4(xr + vi * 8), v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
  }
  const float mu = warp_sum(sum) / E;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + i * 32 < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mu; sq += d * d; }
    }
  }
  const float rs = rsqrtf(warp_sum(sq) / E + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
  bf16* yr = y + static_cast<long long>(row) * E;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8) + 1);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mu) * rs * g[j] + b[j];
      *reinterpret_cast<uint4*>(yr + vi * 8) = pack8(o);
    }
  }
}

int layernorm_fwd(const bf16* x, const float* gamma, const float* beta, bf16* y, float* mean,
                  float* rstd, int rows, int E, float eps, cudaStream_t s) {
  B200_REQUIRE(E % 8 == 0 && E <= LN_MAXV * 256, "layernorm: E (%d) must be a multiple of 8 and <= %d", E, LN_MAXV * 256);
  if (rows == 0) return 0;
  const int nv = cdiv(E / 8, 32);
#define B200_LN_FWD(NV) B200_CHECK_CUDA(launch_kernel(layernorm_fwd_kernel<NV>, dim3(cdiv(rows, 4)), dim3(128), 0, s, true, 1, x, gamma, beta, y, mean, rstd, rows, E, eps))
  if (nv <= 1) B200_LN_FWD(1); else if (nv == 2) B200_LN_FWD(2); else if (nv == 3) B200_LN_FWD(3);
  else if (nv == 4) B200_LN_FWD(4); else B200_LN_FWD(8);
#undef B200_LN_FWD
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// LayerNorm over y = sum_s parts[s] + bias + residual: the consumer of a split-K "partials" GEMM
// (gemm.cuh).  The pre-norm sum never exists in bf16, the bias / residual epilogue of the GEMM
// and the separate LayerNorm pass over its output collapse into this one kernel.
template <int NV>
__global__ void __launch_bounds__(128)
layernorm_reduce_fwd_kernel(const float* __restrict__ parts, int nsplit, long long slab_stride, long long ldp,
                            const float* __restrict__ bias, const bf16* __restrict__ residual, long long ldr,
                            const float* __restrict__ gamma, const float* __restrict__ beta,
                            bf16* __restrict__ y, int rows, int E, float eps) {
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= rows) return;
  const int nvec = E >> 3;
  float v[NV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      if (residual) unpack8(ldg_nc_v4(residual + static_cast<long long>(row) * ldr + vi * 8), v[i]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
      }
      if (bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + vi * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + vi * 8) + 1);
        v[i][0] += b0.x; v[i][1] += b0.y; v[i][2] += b0.z; v[i][3] += b0.w;
        v[i][4] += b1.x; v[i][5] += b1.y; v[i][6] += b1.z; v[i][7] += b1.w;
      }
      const float* pr = parts + static_cast<long long>(row) * ldp + vi * 8;
      for (int s0 = 0; s0 < nsplit; s0 += 4) {     // four slabs (eight 16-byte loads) in flight at a time
        float4 a0[4], a1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          a1[u] = a0[u];
          if (s0 + u < nsplit) {
            a0[u] = *reinterpret_cast<const float4*>(pr + (s0 + u) * slab_stride);
            a1[u] = *(reinterpret_cast<const float4*>(pr + (s0 + u) * slab_stride) + 1);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[i][0] += a0[u].x; v[i][1] += a0[u].y; v[i][2] += a0[u].z; v[i][3] += a0[u].w;
          v[i][4] += a1[u].x; v[i][5] += a1[u].y; v[i][6] += a1[u].z; v[i][7] += a1[u].w;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
  }
  // gamma / beta requested before the two warp reductions so that their latency hides behind them
  float4 gq[NV][2], bq[NV][2];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      gq[i][0] = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      gq[i][1] = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8) + 1);
      bq[i][0] = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
      bq[i][1] = __ldg(reinterpret_cast<const float4*>(beta + vi * 8) + 1);
    }
  }
  const float mu = warp_sum(sum) / E;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + i * 32 < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mu; sq += d * d; }
    }
  }
  const float rs = rsqrtf(warp_sum(sq) / E + eps);
  bf16* yr = y + static_cast<long long>(row) * E;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const float g[8] = {gq[i][0].x, gq[i][0].y, gq[i][0].z, gq[i][0].w, gq[i][1].x, gq[i][1].y, gq[i][1].z, gq[i][1].w};
      const float b[8] = {bq[i][0].x, bq[i][0].y, bq[i][0].z, bq[i][0].w, bq[i][1].x, bq[i][1].y, bq[i][1].z, bq[i][1].w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mu) * rs * g[j] + b[j];
      *reinterpret_cast<uint4*>(yr + vi * 8) = pack8(o);
    }
  }
}

int layernorm_reduce_fwd(const float* parts, int nsplit, long long slab_stride, long long ldp, const float* bias,
                         const bf16* residual, long long ldr, const float* gamma, const float* beta, bf16* y,
                         int rows, int E, float eps, cudaStream_t s) {
  B200_REQUIRE(E % 8 == 0 && E <= LN_MAXV * 256, "layernorm_reduce: E (%d) must be a multiple of 8 and <= %d", E, LN_MAXV * 256);
  B200_REQUIRE(nsplit >= 1 && ldp % 4 == 0 && slab_stride % 4 == 0, "layernorm_reduce: bad partial layout");
  if (rows == 0) return 0;
  const int nv = cdiv(E / 8, 32);
#define B200_LNR(NV) B200_CHECK_CUDA(launch_kernel(layernorm_reduce_fwd_kernel<NV>, dim3(cdiv(rows, 4)), dim3(128), 0, s, true, 1, \
                                                   parts, nsplit, slab_stride, ldp, bias, residual, ldr, gamma, beta, y, rows, E, eps))
  if (nv <= 1) B200_LNR(1); else if (nv == 2) B200_LNR(2); else if (nv == 3) B200_LNR(3);
  else if (nv == 4) B200_LNR(4); else B200_LNR(8);
#undef B200_LNR
  note_launch();
  return 0;
}

// dx = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma;  dgamma += dy*xhat, dbeta += dy,
// and optionally dxsum += dx (the bias gradient of the Linear whose output fed this LayerNorm, so
// that no separate column-sum pass over dx is needed).  Each warp walks rows with a grid stride,
// two rows per iteration (both rows' loads in flight together), keeping its column slice of the
// three column reductions in registers; the 4 warps of a block combine through shared memory and
// issue one atomic per column.
template <int NV>
__global__ void __launch_bounds__(128)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                     const float* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ rstd, bf16* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum,
                     int rows, int E, bf16* __restrict__ dx_drop, const DropCfg dc) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm_red[];  // [3][4][E]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = E >> 3;
  float gam[NV][8], dg[NV][8], db[NV][8], dsx[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
#pragma unroll
    for (int j = 0; j < 8; ++j) { dg[i][j] = 0.f; db[i][j] = 0.f; dsx[i][j] = 0.f; gam[i][j] = 0.f; }
    if (vi < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8) + 1);
      gam[i][0] = g0.x; gam[i][1] = g0.y; gam[i][2] = g0.z; gam[i][3] = g0.w;
      gam[i][4] = g1.x; gam[i][5] = g1.y; gam[i][6] = g1.z; gam[i][7] = g1.w;
    }
  }
  const uint32_t dkey = dc.thr ? drop_key(dc) : 0u;
  const int stride = gridDim.x * 4;
  for (int row0 = blockIdx.x * 4 + warp; row0 < rows; row0 += 2 * stride) {
    uint4 xr[2][NV], dr[2][NV];
    float mu[2], rs[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int row = row0 + q * stride;
      const bool ok = row < rows;
      mu[q] = ok ? mean[row] : 0.f;
      rs[q] = ok ? rstd[row] : 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + i * 32;
        xr[q][i] = make_uint4(0u, 0u, 0u, 0u);
        dr[q][i] = make_uint4(0u, 0u, 0u, 0u);
        if (ok && vi < nvec) {
          xr[q][i] = ldg_nc_v4(x + static_cast<long long>(row) * E + vi * 8);
          dr[q][i] = ldg_nc_v4(dy + static_cast<long long>(row) * E + vi * 8);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int row = row0 + q * stride;
      if (row >= rows) break;   // warp-uniform
      float xh[NV][8], g[NV][8];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + i * 32 < nvec) {
          float xv[8], dv[8];
          unpack8(xr[q][i], xv);
          unpack8(dr[q][i], dv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            xh[i][j] = (xv[j] - mu[q]) * rs[q];
            g[i][j] = dv[j] * gam[i][j];
            s1 += g[i][j];
            s2 += g[i][j] * xh[i][j];
            dg[i][j] += dv[j] * xh[i][j];
            db[i][j] += dv[j];
          }
        }
      }
      const float c1 = warp_sum(s1) / E, c2 = warp_sum(s2) / E;
      bf16* dxr = dx + static_cast<long long>(row) * E;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rs[q] * (g[i][j] - c1 - xh[i][j] * c2);
          *reinterpret_cast<uint4*>(dxr + vi * 8) = pack8(o);
          if (dc.thr) {
            // the Linear that fed this LayerNorm went through dropout (x + dropout(z)): its operand
            // gradient dz = dx * mask / (1 - p) goes to dx_drop, the residual branch keeps dx
            const uint32_t pair0 = (static_cast<uint32_t>(row) * E + vi * 8) >> 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) drop_apply2(o[2 * j], o[2 * j + 1], drop_rand(dkey, pair0 + j), dc.thr, dc.scale);
            *reinterpret_cast<uint4*>(dx_drop + static_cast<long long>(row) * E + vi * 8) = pack8(o);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) dsx[i][j] += o[j];
        }
      }
    }
  }
  float* sg = sm_red;
  float* sb = sm_red + 4 * E;
  float* sx = sm_red + 8 * E;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sg[warp * E + vi * 8 + j] = dg[i][j];
        sb[warp * E + vi * 8 + j] = db[i][j];
        sx[warp * E + vi * 8 + j] = dsx[i][j];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    atomicAdd(dgamma + c, sg[c] + sg[E + c] + sg[2 * E + c] + sg[3 * E + c]);
    atomicAdd(dbeta + c, sb[c] + sb[E + c] + sb[2 * E + c] + sb[3 * E + c]);
    if (dxsum) atomicAdd(dxsum + c, sx[c] + sx[E + c] + sx[2 * E + c] + sx[3 * E + c]);
  }
}

int layernorm_bwd(const bf16* dy, const bf16* x, const float* gamma, const float* mean,
                  const float* rstd, bf16* dx, float* dgamma, float* dbeta, float* dxsum, int rows, int E,
                  cudaStream_t s, bf16* dx_drop, DropCfg dc) {
  B200_REQUIRE(dc.thr == 0 || dx_drop != nullptr, "layernorm_bwd: dropout needs the second output");
  B200_REQUIRE(E % 8 == 0 && E <= LN_MAXV * 256, "layernorm_bwd: E (%d) must be a multiple of 8 and <= %d", E, LN_MAXV * 256);
  if (rows == 0) return 0;
  int blocks = cdiv(rows, 8);          // two rows per warp iteration
  const int cap = 2 * 148;             // measured on B200 (M=12032, E=768): 2 CTAs/SM 26.5 us, 2.5: 34.8, 3: 31.7, 4: 28.7 us, 1: 32.8 us;
                                       // the column-sum atomics at the end cost ~2 us of it (measured by leaving them out)
  if (blocks > cap) blocks = cap;
  const size_t smem = static_cast<size_t>(12) * E * sizeof(float);
  const int nv = cdiv(E / 8, 32);
#define B200_LN_BWD(NV)                                                                              \
  do {                                                                                               \
    static bool configured = false;                                                                  \
    if (!configured) {                                                                               \
      B200_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * NV * 256 * 4)); \
      configured = true;                                                                             \
    }                                                                                                \
    B200_CHECK_CUDA(launch_kernel(layernorm_bwd_kernel<NV>, dim3(blocks), dim3(128), smem, s, true, 1, dy, x, gamma, mean, rstd, dx, dgamma, dbeta, dxsum, rows, E, dx_drop, dc)); \
  } while (0)
  if (nv <= 1) B200_LN_BWD(1); else if (nv == 2) B200_LN_BWD(2); else if (nv == 3) B200_LN_BWD(3);
  else if (nv == 4) B200_LN_BWD(4); else B200_LN_BWD(8);
#undef B200_LN_BWD
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// column sums (bias gradients): out[n] += sum_m x[m,n]
// ------------------------------------------------------------------------------------------
// The training backward runs these passes on a side stream next to the persistent GEMM CTAs, which give up one
// TMA stage for them (24-48 KB): 8 KB of shared memory and <= 40 registers per thread fit beside a GEMM CTA.
// Block = 256 columns x `rows_per_cta` rows (a multiple of 64): warp = row lane, its 32 lanes cover one 512-byte row
// segment; eight 16-byte loads in flight per thread; the eight row lanes fold through shared memory and every
// column gets ONE atomic per CTA.  Same-address atomics serialise in L2 (a 50432-row matrix summed in 64-row
// blocks spent 160 us on 1576 atomics per column), so the host sizes the row extent for ~3 CTAs per SM.
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ x, long long ldx, float* __restrict__ out, int M, int N, int rows_per_cta) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8][256];
  const int cv = threadIdx.x & 31;   // column vector within the block's 256-column strip
  const int rl = threadIdx.x >> 5;   // row lane 0..7
  const int col = blockIdx.x * 256 + cv * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < N) {
    const bf16* base = x + col;
    const int r_begin = blockIdx.y * rows_per_cta;
    const int r_end = min(r_begin + rows_per_cta, M);
    for (int r0 = r_begin + rl; r0 < r_end; r0 += 64) {
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rr = r0 + 8 * k;
        u[k] = (rr < r_end) ? ldg_nc_v4(base + static_cast<long long>(rr) * ldx) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cv * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < N) {
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) sum += red[r][c];
    atomicAdd(out + blockIdx.x * 256 + c, sum);
  }
}

int colsum(const bf16* x, long long ldx, float* out, int M, int N, cudaStream_t s) {
  B200_REQUIRE(N % 8 == 0 && ldx % 8 == 0, "colsum: N (%d) and ldx (%lld) must be multiples of 8", N, ldx);
  if (M == 0) return 0;
  const int col_blocks = cdiv(N, 256);
  // ~3 CTAs per SM over the whole grid, in whole 64-row slabs
  const int want_row_blocks = max(1, (3 * 148) / col_blocks);
  int rows_per_cta = cdiv(cdiv(M, want_row_blocks), 64) * 64;
  if (rows_per_cta < 64) rows_per_cta = 64;
  dim3 grid(col_blocks, cdiv(M, rows_per_cta));
  B200_CHECK_CUDA(launch_kernel(colsum_kernel, grid, dim3(256), 0, s, true, 1, x, ldx, out, M, N, rows_per_cta));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// casts
// ------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + i) + 1);
    uint4 o;
    o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w);
    o.z = pack_bf16(b.x, b.y); o.w = pack_bf16(b.z, b.w);
    *reinterpret_cast<uint4*>(dst + i) = o;
  } else {
    for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
  }
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(src + i), f);
    *reinterpret_cast<float4*>(dst + i) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(dst + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    for (long long j = i; j < n; ++j) dst[j] = __bfloat162float(src[j]);
  }
}
int cast_f32_to_bf16(const float* src, bf16* dst, long long n, cudaStream_t s) {
  if (n == 0) return 0;
  B200_REQUIRE(((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0), "cast: pointers must be 16-byte aligned");
  B200_CHECK_CUDA(launch_kernel(cast_f32_bf16_kernel, dim3(cdiv(cdiv(n, 8), 256)), dim3(256), 0, s, true, 1, src, dst, n));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int cast_bf16_to_f32(const bf16* src, float* dst, long long n, cudaStream_t s) {
  if (n == 0) return 0;
  B200_REQUIRE(((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0), "cast: pointers must be 16-byte aligned");
  B200_CHECK_CUDA(launch_kernel(cast_bf16_f32_kernel, dim3(cdiv(cdiv(n, 8), 256)), dim3(256), 0, s, true, 1, src, dst, n));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// softmax-CE finalisation
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ce_finalize_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                   const float* __restrict__ tgt_logit, const int64_t* __restrict__ targets, int M,
                   int n_tiles, long long ignore_index, float* __restrict__ row_lse,
                   float* __restrict__ row_loss, float* __restrict__ loss_sum,
                   float* __restrict__ valid_count) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.f, cnt = 0.f;
  if (row < M) {
    const float* pm = part_max + static_cast<long long>(row) * n_tiles;
    const float* ps = part_sum + static_cast<long long>(row) * n_tiles;
    float m = -INFINITY;
    for (int i = 0; i < n_tiles; ++i) m = fmaxf(m, pm[i]);
    float s = 0.f;
    for (int i = 0; i < n_tiles; ++i) s += ps[i] * expf(pm[i] - m);
    const float lse = m + logf(s);
    row_lse[row] = lse;
    const long long t = targets[row];
    if (t != ignore_index) {
      loss = lse - tgt_logit[row];
      cnt = 1.f;
    }
    if (row_loss) row_loss[row] = loss;
  }
  loss = warp_sum(loss);
  cnt = warp_sum(cnt);
  __shared__ float sl[8], sc[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sl[warp] = loss; sc[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) { a += sl[i]; b += sc[i]; }
    atomicAdd(loss_sum, a);
    atomicAdd(valid_count, b);
  }
}

int ce_finalize(const float* part_max, const float* part_sum, const float* tgt_logit,
                const int64_t* targets, int M, int n_tiles, long long ignore_index, float* row_lse,
                float* row_loss, float* loss_sum, float* valid_count, cudaStream_t s) {
  B200_CHECK_CUDA(launch_kernel(ce_finalize_kernel, dim3(cdiv(M, 256)), dim3(256), 0, s, true, 1, part_max, part_sum, tgt_logit, targets, M, n_tiles,
                                                 ignore_index, row_lse, row_loss, loss_sum, valid_count));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void ce_mean_kernel(const float* loss_sum, const float* valid_count, float* out) {
  pdl_wait();
  pdl_trigger();
  const float c = *valid_count;
  out[0] = *loss_sum / c;   // 0/0 -> NaN, as nn.CrossEntropyLoss does when every target is ignored
  out[1] = c;
  out[2] = 1.f / c;
}
int ce_mean(const float* loss_sum, const float* valid_count, float* out, cudaStream_t s) {
  B200_CHECK_CUDA(launch_kernel(ce_mean_kernel, dim3(1), dim3(1), 0, s, true, 1, loss_sum, valid_count, out));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// one warp per row: lanes scan the per-tile partials, then a shuffle reduction keeps the maximum
// with the LOWEST column index on ties (torch.argmax returns the first maximal index)
__global__ void __launch_bounds__(256)
argmax_finalize_kernel(const float* __restrict__ part_max, const float* __restrict__ part_idx,
                       int M, int n_tiles, int64_t* __restrict__ out_ids, float* __restrict__ out_max) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* pm = part_max + static_cast<long long>(row) * n_tiles;
  const float* pi = part_idx + static_cast<long long>(row) * n_tiles;
  float best = -INFINITY;
  int idx = 0x7fffffff;
  for (int i = lane; i < n_tiles; i += 32) {
    const float v = pm[i];
    const int id = __float_as_int(pi[i]);
    if (v > best || (v == best && id < idx)) { best = v; idx = id; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
  }
  if (lane == 0) {
    out_ids[row] = idx == 0x7fffffff ? 0 : idx;
    if (out_max) out_max[row] = best;
  }
}
int argmax_finalize(const float* part_max, const float* part_idx, int M, int n_tiles, int64_t* out_ids,
                    float* out_max, cudaStream_t s) {
  B200_CHECK_CUDA(launch_kernel(argmax_finalize_kernel, dim3(cdiv(M, 8)), dim3(256), 0, s, true, 1, part_max, part_idx, M, n_tiles, out_ids, out_max));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// optimizer: global grad norm + clip + AdamW (train.py:96-100, 319-325)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ sumsq) {
  pdl_wait();
  pdl_trigger();
  float acc = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g + i));
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (long long j = i; j < n; ++j) acc += g[j] * g[j];
    }
  }
  acc = warp_sum(acc);
  __shared__ float sw[8];
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sw[i];
    atomicAdd(sumsq, t);
  }
}
int grad_sumsq(const float* g, long long n, float* sumsq, cudaStream_t s) {
  if (n == 0) return 0;
  int blocks = cdiv(n, 256 * 4 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  B200_CHECK_CUDA(launch_kernel(grad_sumsq_kernel, dim3(blocks), dim3(256), 0, s, true, 1, g, n, sumsq));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, bf16* __restrict__ p16, const float* __restrict__ g,
             float* __restrict__ m, float* __restrict__ v, long long n, const float* __restrict__ sumsq,
             float max_norm, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
             const int* __restrict__ step_dev, const float* __restrict__ lr_dev) {
  pdl_wait();
  pdl_trigger();
  if (step_dev != nullptr) {   // graph-replay mode: the step count (and lr) live on the device
    const float t = static_cast<float>(*step_dev);
    bc1 = 1.f - powf(b1, t);
    bc2_sqrt = sqrtf(1.f - powf(b2, t));
  }
  if (lr_dev != nullptr) lr = *lr_dev;
  float coef = 1.f;
  if (max_norm > 0.f && sumsq != nullptr) {
    const float total = sqrtf(*sumsq);
    coef = fminf(1.f, max_norm / (total + 1e-6f));
  }
  const float step_size = lr / bc1;
  const long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4;
  if (i >= n) return;
  float pv[4], gv[4], mv[4], vv[4];
  const int cnt = (i + 4 <= n) ? 4 : static_cast<int>(n - i);
  if (cnt == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p + i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(g + i));
    const float4 c = *reinterpret_cast<const float4*>(m + i);
    const float4 d = *reinterpret_cast<const float4*>(v + i);
    pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w;
    gv[0] = b.x; gv[1] = b.y; gv[2] = b.z; gv[3] = b.w;
    mv[0] = c.x; mv[1] = c.y; mv[2] = c.z; mv[3] = c.w;
    vv[0] = d.x; vv[1] = d.y; vv[2] = d.z; vv[3] = d.w;
  } else {
    for (int j = 0; j < 4; ++j) {
      const bool ok = j < cnt;
      pv[j] = ok ? p[i + j] : 0.f; gv[j] = ok ? g[i + j] : 0.f;
      mv[j] = ok ? m[i + j] : 0.f; vv[j] = ok ? v[i + j] : 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float gg = gv[j] * coef;
    pv[j] *= (1.f - lr * wd);
    mv[j] = b1 * mv[j] + (1.f - b1) * gg;
    vv[j] = b2 * vv[j] + (1.f - b2) * gg * gg;
    const float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
    pv[j] -= step_size * (mv[j] / denom);
  }
  if (cnt == 4) {
    *reinterpret_cast<float4*>(p + i) = make_float4(pv[0], pv[1], pv[2], pv[3]);
    *reinterpret_cast<float4*>(m + i) = make_float4(mv[0], mv[1], mv[2], mv[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (p16) {
      uint2 o;
      o.x = pack_bf16(pv[0], pv[1]); o.y = pack_bf16(pv[2], pv[3]);
      *reinterpret_cast<uint2*>(p16 + i) = o;
    }
  } else {
    for (int j = 0; j < cnt; ++j) {
      p[i + j] = pv[j]; m[i + j] = mv[j]; v[i + j] = vv[j];
      if (p16) p16[i + j] = __float2bfloat16(pv[j]);
    }
  }
}

__global__ void inc_step_kernel(int* step) {
  pdl_wait();
  pdl_trigger(); *step += 1; }

int adamw_step(float* p, bf16* p16, const float* g, float* m, float* v, long long n,
               const float* sumsq, float max_norm, float lr, float b1, float b2, float eps, float wd,
               int step, cudaStream_t s, int* step_dev, const float* lr_dev) {
  if (n == 0) return 0;
  B200_REQUIRE(step >= 1 || step_dev != nullptr, "adamw: step must be >= 1");
  if (step_dev != nullptr) {
    B200_CHECK_CUDA(launch_kernel(inc_step_kernel, dim3(1), dim3(1), 0, s, true, 1, step_dev));
    note_launch();
    if (step < 1) step = 1;
  }
  const float bc1 = 1.f - powf(b1, static_cast<float>(step));
  const float bc2 = 1.f - powf(b2, static_cast<float>(step));
  const double bc1d = 1.0 - pow(static_cast<double>(b1), step);
  const double bc2d = 1.0 - pow(static_cast<double>(b2), step);
  (void)bc1; (void)bc2;
  B200_CHECK_CUDA(launch_kernel(adamw_kernel, dim3(cdiv(cdiv(n, 4), 256)), dim3(256), 0, s, true, 1, p, p16, g, m, v, n, sumsq, max_norm, lr, b1, b2, eps,
                                                      wd, static_cast<float>(bc1d),
                                                      static_cast<float>(sqrt(bc2d)), step_dev, lr_dev));
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
