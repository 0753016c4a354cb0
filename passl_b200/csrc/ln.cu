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

// CTA = LN_BWD_WARPS warps, each warp walks rows r0 + w, r0 + w + WARPS, ... of the CTA's row range.
// part[blockIdx.x][0][D] = sum_rows dy (dbeta), part[blockIdx.x][1][D] = sum_rows dy * xhat (dgamma)
// Round 2: the pass was latency-bound (2.2-3.0 TB/s, 11.5 % of the CLIP step): one row per warp at a time with ~200 registers
// of fp32 row state kept only 37 KB per SM in flight.  Now the row stays PACKED (bf16, 4 registers per 8 elements) and is
// unpacked twice, the next row of the warp is requested before the current one is touched, and 12 warps share an SM.
constexpr int LN_BWD_WARPS = 12;
template <int MAXCH>
__global__ void __launch_bounds__(LN_BWD_WARPS * 32) ln_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                                   const float* __restrict__ gamma, const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ dres,
                                                                   __nv_bfloat16* __restrict__ dx, float* __restrict__ part, long long T,
                                                                   int D, int rows_per_block) {
  extern __shared__ float red[];  // [D] gamma, then [WARPS][2][D] partials
  float* gsm = red;
  float* rsm = red + D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = D >> 3;
  for (int i = threadIdx.x; i < D; i += blockDim.x) gsm[i] = gamma[i];
  __syncthreads();
  float ab[MAXCH][8], ag[MAXCH][8];
#pragma unroll
  for (int j = 0; j < MAXCH; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) { ab[j][e] = 0.f; ag[j][e] = 0.f; }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > T) r1 = T;
  uint4 px[MAXCH], pd[MAXCH], pr[MAXCH];
  auto load_row = [&](long long row, uint4 (&ax)[MAXCH], uint4 (&ad)[MAXCH], uint4 (&ar)[MAXCH]) {
#pragma unroll
    for (int j = 0; j < MAXCH; ++j) {
      const int c = lane + 32 * j;
      if (c < chunks) {
        ax[j] = ld_nc_v4(x + row * D + c * 8);
        ad[j] = ld_nc_v4(dy + row * D + c * 8);
        if (dres) ar[j] = ld_nc_v4(dres + row * D + c * 8);
      }
    }
  };
  long long row = r0 + warp;
  float mu = 0.f, rs = 0.f;
  if (row < r1) { load_row(row, px, pd, pr); mu = mean[row]; rs = rstd[row]; }
  for (; row < r1; row += LN_BWD_WARPS) {
    uint4 nx[MAXCH], nd[MAXCH], nr[MAXCH];
    const long long nrow = row + LN_BWD_WARPS;
    float nmu = 0.f, nrs = 0.f;
    if (nrow < r1) { load_row(nrow, nx, nd, nr); nmu = mean[nrow]; nrs = rstd[nrow]; }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < MAXCH; ++j) {
      const int c = lane + 32 * j;
      if (c < chunks) {
        float xv[8], dv[8];
        ln_unpack8(px[j], xv);
        ln_unpack8(pd[j], dv);
        const float4 g0 = *reinterpret_cast<const float4*>(gsm + c * 8), g1 = *reinterpret_cast<const float4*>(gsm + c * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (xv[e] - mu) * rs;
          const float g = dv[e] * gg[e];
          s1 += g;
          s2 = fmaf(g, xh, s2);
          ab[j][e] += dv[e];
          ag[j][e] = fmaf(dv[e], xh, ag[j][e]);
        }
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int j = 0; j < MAXCH; ++j) {
      const int c = lane + 32 * j;
      if (c < chunks) {
        float xv[8], dv[8], o[8];
        ln_unpack8(px[j], xv);
        ln_unpack8(pd[j], dv);
        const float4 g0 = *reinterpret_cast<const float4*>(gsm + c * 8), g1 = *reinterpret_cast<const float4*>(gsm + c * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = rs * (dv[e] * gg[e] - s1 - (xv[e] - mu) * rs * s2);
        if (dres) {   // gradient arriving through the residual connection around this LayerNorm's branch
          float rr[8];
          ln_unpack8(pr[j], rr);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] += rr[e];
        }
        *reinterpret_cast<uint4*>(dx + row * D + c * 8) = ln_pack8(o);
      }
    }
#pragma unroll
    for (int j = 0; j < MAXCH; ++j) { px[j] = nx[j]; pd[j] = nd[j]; pr[j] = nr[j]; }
    mu = nmu; rs = nrs;
  }
  // cross-warp reduction of the parameter-gradient partials
#pragma unroll
  for (int j = 0; j < MAXCH; ++j) {
    const int c = lane + 32 * j;
    if (c < chunks) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        rsm[(warp * 2 + 0) * D + c * 8 + e] = ab[j][e];
        rsm[(warp * 2 + 1) * D + c * 8 + e] = ag[j][e];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
    const int which = i / D, col = i - which * D;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_BWD_WARPS; ++w) s += rsm[(w * 2 + which) * D + col];
    part[((size_t)blockIdx.x * 2 + which) * D + col] = s;
  }
}

static void ln_bwd_cfg(long long T, int& blocks, int& rpb) {
  long long target = (long long)num_sms();        // one 12-warp CTA per SM
  long long r = (T + target - 1) / target;
  if (r < LN_BWD_WARPS) r = LN_BWD_WARPS;
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
  const int smem = (LN_BWD_WARPS * 2 + 1) * D * 4;
  static bool attr = false;
  if (!attr) {
    PB_CUDA_CHECK(cudaFuncSetAttribute(ln_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (LN_BWD_WARPS * 2 + 1) * 512 * 4));
    PB_CUDA_CHECK(cudaFuncSetAttribute(ln_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (LN_BWD_WARPS * 2 + 1) * 768 * 4));
    PB_CUDA_CHECK(cudaFuncSetAttribute(ln_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (LN_BWD_WARPS * 2 + 1) * 1024 * 4));
    PB_CUDA_CHECK(cudaFuncSetAttribute(ln_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (LN_BWD_WARPS * 2 + 1) * 2048 * 4));
    attr = true;
  }
#define PB_LN_BWD(mc)                                                                                                            \
  ln_bwd_kernel<mc><<<blocks, LN_BWD_WARPS * 32, smem, (cudaStream_t)stream>>>(                                                  \
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), gamma, mean, rstd,                  \
      reinterpret_cast<const __nv_bfloat16*>(dres), reinterpret_cast<__nv_bfloat16*>(dx), part, T, D, rpb)
  if (D <= 512) PB_LN_BWD(2);
  else if (D <= 768) PB_LN_BWD(3);
  else if (D <= 1024) PB_LN_BWD(4);
  else PB_LN_BWD(8);
#undef PB_LN_BWD
  PB_LAUNCH_CHECK();
  return PB_OK;
}
