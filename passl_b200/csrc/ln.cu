// LayerNorm forward / backward over token rows [T, D] bf16 (fp32 statistics) — ViT / MAE blocks.
// Replaces paddle nn.LayerNorm at passl/models/vision_transformer.py:174,204-205 (eps 1e-6: vision_transformer.py:440,
// mae.py:52-53).  HBM-bound: forward reads x once and writes y once (re-reads hit L1); backward reads x, dy once, writes dx
// once; dgamma / dbeta are produced as per-CTA partials (no atomics) summed by passl_b200_bn_bwd_finalize.
#include "common.cuh"
#include "host_utils.h"
#include "../../include/passl_b200.h"

namespace pb {

__device__ __forceinline__ void ln_unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 ln_pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

constexpr int LN_MAXCH = 8;  // 8-element chunks per lane -> D <= 2048 (MAXCH = 4 instantiation for D <= 1024)

// one warp per row
template <int MAXCH>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd, long long T, int D,
                                                     float eps) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= T) return;
  const int chunks = D >> 3;
  const __nv_bfloat16* xr = x + row * D;
  float v[MAXCH][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < MAXCH; ++j) {
    const int c = lane + 32 * j;
    if (c < chunks) {
      ln_unpack8(*reinterpret_cast<const uint4*>(xr + c * 8), v[j]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[j][e];
    }
  }
  const float mu = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < MAXCH; ++j) {
    const int c = lane + 32 * j;
    if (c < chunks) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[j][e] - mu; q = fmaf(d, d, q); }
    }
  }
  const float rs = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
  for (int j = 0; j < MAXCH; ++j) {
    const int c = lane + 32 * j;
    if (c < chunks) {
      float o[8];
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + c * 8), g1 = *reinterpret_cast<const float4*>(gamma + c * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + c * 8), b1 = *reinterpret_cast<const float4*>(beta + c * 8 + 4);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[j][e] - mu) * rs * gg[e] + bb[e];
      *reinterpret_cast<uint4*>(y + row * D + c * 8) = ln_pack8(o);
    }
  }
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
}

// CTA = 8 warps, each warp walks rows r0 + w, r0 + w + 8, ... of the CTA's row range.
// part[blockIdx.x][0][D] = sum_rows dy (dbeta), part[blockIdx.x][1][D] = sum_rows dy * xhat (dgamma)
template <int MAXCH>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ dres,
                                                     __nv_bfloat16* __restrict__ dx, float* __restrict__ part, long long T,
                                                     int D, int rows_per_block) {
  extern __shared__ float red[];  // [8 warps][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = D >> 3;
  float ab[MAXCH][8], ag[MAXCH][8];
#pragma unroll
  for (int j = 0; j < MAXCH; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) { ab[j][e] = 0.f; ag[j][e] = 0.f; }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > T) r1 = T;
  for (long long row = r0 + warp; row < r1; row += 8) {
    const float mu = mean[row], rs = rstd[row];
    float xh[MAXCH][8], g[MAXCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < MAXCH; ++j) {
      const int c = lane + 32 * j;
      if (c < chunks) {
        float xv[8], dv[8];
        ln_unpack8(ld_nc_v4(x + row * D + c * 8), xv);
        ln_unpack8(ld_nc_v4(dy + row * D + c * 8), dv);
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + c * 8), g1 = *reinterpret_cast<const float4*>(gamma + c * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          xh[j][e] = (xv[e] - mu) * rs;
          g[j][e] = dv[e] * gg[e];
          s1 += g[j][e];
          s2 = fmaf(g[j][e], xh[j][e], s2);
          ab[j][e] += dv[e];
          ag[j][e] = fmaf(dv[e], xh[j][e], ag[j][e]);
        }
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int j = 0; j < MAXCH; ++j) {
      const int c = lane + 32 * j;
      if (c < chunks) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = rs * (g[j][e] - s1 - xh[j][e] * s2);
        if (dres) {   // gradient arriving through the residual connection around this LayerNorm's branch
          float rr[8];
          ln_unpack8(ld_nc_v4(dres + row * D + c * 8), rr);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] += rr[e];
        }
        *reinterpret_cast<uint4*>(dx + row * D + c * 8) = ln_pack8(o);
      }
    }
  }
  // cross-warp reduction of the parameter-gradient partials
#pragma unroll
  for (int j = 0; j < MAXCH; ++j) {
    const int c = lane + 32 * j;
    if (c < chunks) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        red[(warp * 2 + 0) * D + c * 8 + e] = ab[j][e];
        red[(warp * 2 + 1) * D + c * 8 + e] = ag[j][e];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
    const int which = i / D, col = i - which * D;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[(w * 2 + which) * D + col];
    part[((size_t)blockIdx.x * 2 + which) * D + col] = s;
  }
}

static void ln_bwd_cfg(long long T, int& blocks, int& rpb) {
  long long target = (long long)num_sms() * 2;
  long long r = (T + target - 1) / target;
  if (r < 8) r = 8;
  rpb = (int)r;
  blocks = (int)((T + r - 1) / r);
}

}  // namespace pb

using namespace pb;

extern "C" int passl_b200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                        float* rstd, long long T, int D, float eps, void* stream) {
  if (T <= 0 || D <= 0 || D % 8 || D > 256 * LN_MAXCH) return PB_ERR_BAD_ARG;
  if (D <= 1024)
    ln_fwd_kernel<4><<<(unsigned)((T + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, reinterpret_cast<__nv_bfloat16*>(y), mean, rstd, T, D, eps);
  else
    ln_fwd_kernel<8><<<(unsigned)((T + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, reinterpret_cast<__nv_bfloat16*>(y), mean, rstd, T, D, eps);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_layernorm_bwd_blocks(long long T) {
  int b, r;
  ln_bwd_cfg(T, b, r);
  return b;
}

// part: fp32 [nblk, 2, D] partials (dbeta, dgamma) -> sum with passl_b200_bn_bwd_finalize
extern "C" int passl_b200_layernorm_bwd(const void* x, const void* dy, const float* gamma, const float* mean,
                                        const float* rstd, const void* dres, void* dx, float* part, long long T, int D,
                                        void* stream) {
  if (T <= 0 || D <= 0 || D % 8 || D > 256 * LN_MAXCH) return PB_ERR_BAD_ARG;
  int blocks, rpb;
  ln_bwd_cfg(T, blocks, rpb);
  const int smem = 8 * 2 * D * 4;
  static bool attr = false;
  if (!attr) {
    PB_CUDA_CHECK(cudaFuncSetAttribute(ln_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 1024 * 4));
    PB_CUDA_CHECK(cudaFuncSetAttribute(ln_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 2048 * 4));
    attr = true;
  }
  if (D <= 1024)
    ln_bwd_kernel<4><<<blocks, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                 reinterpret_cast<const __nv_bfloat16*>(dy), gamma, mean, rstd,
                                                                 reinterpret_cast<const __nv_bfloat16*>(dres),
                                                                 reinterpret_cast<__nv_bfloat16*>(dx), part, T, D, rpb);
  else
    ln_bwd_kernel<8><<<blocks, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                 reinterpret_cast<const __nv_bfloat16*>(dy), gamma, mean, rstd,
                                                                 reinterpret_cast<const __nv_bfloat16*>(dres),
                                                                 reinterpret_cast<__nv_bfloat16*>(dx), part, T, D, rpb);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
