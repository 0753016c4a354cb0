// BatchNorm (training mode) over channels-last activations [P, C] bf16 — statistics, fused apply (+ReLU, +residual)
// and the two-pass backward.  HBM-bound: every pass reads/writes each activation exactly once with 16-byte accesses.
//
// Reference: paddle nn.BatchNorm2D / BatchNorm1D as used at passl_v110/modeling/backbones/resnetimagenet.py:112-131
// (conv-bn-relu, bn3 + identity + relu) and necks/base_neck.py:221-227 (fc-BN1D-ReLU).  Paddle conventions:
// eps 1e-5, momentum 0.9 (running = 0.9*running + 0.1*batch), biased batch variance for normalisation.
#include "common.cuh"
#include "host_utils.h"
#include "../../include/passl_b200.h"

namespace pb {

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// blockDim = (GX, GY): x over 8-channel groups, y over rows.  gridDim = (row blocks, channel-group blocks).
// Each CTA writes ONE partial per channel to part[blockIdx.x][0|1][C] — no atomics (same-address L2 atomics from ~1000 CTAs
// serialise at ~50 ns each and made this kernel 40x slower than its HBM roofline), deterministic summation order.
//   MODE 0 (stats)      : a = y,               second = y*y
//   MODE 1 (bwd reduce) : a = dz*mask,         second = a * (y - mean) * invstd
template <int MODE>
__global__ void __launch_bounds__(256, 2) bn_reduce_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dz,
                                 const __nv_bfloat16* __restrict__ z, const float* __restrict__ mean,
                                 const float* __restrict__ invstd, float* __restrict__ part,
                                 long long P, int C, int rows_per_block, int relu) {
  extern __shared__ float red[];  // [GY][GX*16]
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;  // channel group
  const int c0 = cg * 8;
  const bool cok = c0 < C;
  float s0[8], s1[8], mu[8], is[8], sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s0[i] = 0.f; s1[i] = 0.f; mu[i] = 0.f; is[i] = 1.f; sc[i] = 0.f; sh[i] = 0.f; }
  if (MODE >= 1 && cok) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { mu[i] = mean[c0 + i]; is[i] = invstd[c0 + i]; }
  }
  // relu == 2: the ReLU mask is recomputed from y (z = relu(fma(y, scale, shift)) has no residual term), so z is never read;
  // the `z` argument then carries the fp32 [2, C] (scale, shift) rows of the forward pass
  constexpr bool recompute = (MODE == 2);                 // separate instantiation: the z-reading variant keeps its registers
  const bool read_z = (MODE == 1) && relu == 1;
  constexpr bool read_bits = (MODE == 3);                 // relu == 3: `z` carries the 1-bit-per-element mask written by bn_apply_mask
  const unsigned char* zbits = reinterpret_cast<const unsigned char*>(z);
  const int C8r = C / 8;
  if (recompute && cok) {
    const float* ss = reinterpret_cast<const float*>(z);
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = ss[c0 + i]; sh[i] = ss[C + c0 + i]; }
  }
  const long long r_begin = (long long)blockIdx.x * rows_per_block;
  long long r_end = r_begin + rows_per_block;
  if (r_end > P) r_end = P;
  if (cok) {
    constexpr int U = 4;  // independent 16-byte loads in flight per thread
    long long r = r_begin + threadIdx.y;
    const long long step = blockDim.y;
    for (; r + (U - 1) * step < r_end; r += U * step) {
      uint4 uy[U], ug[U], uz[U];
      unsigned ub[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        uy[u] = ld_nc_v4(y + (r + u * step) * C + c0);
        if (MODE >= 1) {
          ug[u] = ld_nc_v4(dz + (r + u * step) * C + c0);
          if (read_z) uz[u] = ld_nc_v4(z + (r + u * step) * C + c0);
          if (read_bits) ub[u] = zbits[(r + u * step) * C8r + cg];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float a[8];
        unpack8(uy[u], a);
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { s0[i] += a[i]; s1[i] = fmaf(a[i], a[i], s1[i]); }
        } else {
          float g[8];
          unpack8(ug[u], g);
          if (read_z) {
            float zz[8];
            unpack8(uz[u], zz);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = zz[i] > 0.f ? g[i] : 0.f;
          } else if (recompute) {
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = fmaf(a[i], sc[i], sh[i]) > 0.f ? g[i] : 0.f;
          } else if (read_bits) {
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = ((ub[u] >> i) & 1u) ? g[i] : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) { s0[i] += g[i]; s1[i] = fmaf(g[i], (a[i] - mu[i]) * is[i], s1[i]); }
        }
      }
    }
    for (; r < r_end; r += step) {
      float a[8];
      unpack8(ld_nc_v4(y + r * C + c0), a);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { s0[i] += a[i]; s1[i] = fmaf(a[i], a[i], s1[i]); }
      } else {
        float g[8];
        unpack8(ld_nc_v4(dz + r * C + c0), g);
        if (read_z) {
          float zz[8];
          unpack8(ld_nc_v4(z + r * C + c0), zz);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = zz[i] > 0.f ? g[i] : 0.f;
        } else if (recompute) {
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = fmaf(a[i], sc[i], sh[i]) > 0.f ? g[i] : 0.f;
        } else if (read_bits) {
          const unsigned mb = zbits[r * C8r + cg];
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = ((mb >> i) & 1u) ? g[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { s0[i] += g[i]; s1[i] = fmaf(g[i], (a[i] - mu[i]) * is[i], s1[i]); }
      }
    }
  }
  // tree reduction over y in shared memory
  float* my = red + (threadIdx.y * blockDim.x + threadIdx.x) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) { my[i] = s0[i]; my[8 + i] = s1[i]; }
  __syncthreads();
  for (int half = blockDim.y >> 1; half > 0; half >>= 1) {
    if ((int)threadIdx.y < half) {
      const float* o = red + ((threadIdx.y + half) * blockDim.x + threadIdx.x) * 16;
#pragma unroll
      for (int i = 0; i < 16; ++i) my[i] += o[i];
    }
    __syncthreads();
  }
  if (threadIdx.y == 0 && cok) {
    float* dst = part + (size_t)blockIdx.x * 2 * C;
#pragma unroll
    for (int i = 0; i < 8; ++i) { dst[c0 + i] = my[i]; dst[C + c0 + i] = my[8 + i]; }
  }
}

// sums -> mean / invstd / fused scale+shift, running statistics update (Paddle momentum convention).
__global__ void bn_finalize_kernel(const float* __restrict__ part, int nblk,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* mean,
                                   float* invstd, float* scale, float* shift, float* running_mean, float* running_var,
                                   float inv_count, float eps, float momentum, int C) {
  // block (32, 32): x = channel (coalesced partial reads), y = group of partials; smem tree over y
  __shared__ float r0[32][33], r1[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float sm = 0.f, sq = 0.f;
  if (c < C)
    for (int b = threadIdx.y; b < nblk; b += 32) { sm += part[(size_t)b * 2 * C + c]; sq += part[(size_t)b * 2 * C + C + c]; }
  r0[threadIdx.y][threadIdx.x] = sm; r1[threadIdx.y][threadIdx.x] = sq;
  __syncthreads();
  for (int h = 16; h > 0; h >>= 1) {
    if ((int)threadIdx.y < h) { r0[threadIdx.y][threadIdx.x] += r0[threadIdx.y + h][threadIdx.x]; r1[threadIdx.y][threadIdx.x] += r1[threadIdx.y + h][threadIdx.x]; }
    __syncthreads();
  }
  if (threadIdx.y != 0 || c >= C) return;
  sm = r0[0][threadIdx.x]; sq = r1[0][threadIdx.x];
  float m = sm * inv_count;
  float v = fmaxf(sq * inv_count - m * m, 0.f);
  float is = rsqrtf(v + eps);
  float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean[c] = m; invstd[c] = is;
  scale[c] = g * is;
  shift[c] = b - m * g * is;
  if (running_mean) running_mean[c] = momentum * running_mean[c] + (1.f - momentum) * m;
  if (running_var) running_var[c] = momentum * running_var[c] + (1.f - momentum) * v;
}

// z = act(y*scale + shift + residual)
__global__ void bn_apply_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ residual,
                                const float* __restrict__ scale, const float* __restrict__ shift,
                                __nv_bfloat16* __restrict__ z, float* __restrict__ z_f32, long long total8, int C8,
                                int relu, unsigned char* __restrict__ mask) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % C8) * 8;
    float a[8];
    unpack8(ld_nc_v4(y + i * 8), a);
    float4 s0 = *reinterpret_cast<const float4*>(scale + c0), s1 = *reinterpret_cast<const float4*>(scale + c0 + 4);
    float4 h0 = *reinterpret_cast<const float4*>(shift + c0), h1 = *reinterpret_cast<const float4*>(shift + c0 + 4);
    a[0] = fmaf(a[0], s0.x, h0.x); a[1] = fmaf(a[1], s0.y, h0.y); a[2] = fmaf(a[2], s0.z, h0.z); a[3] = fmaf(a[3], s0.w, h0.w);
    a[4] = fmaf(a[4], s1.x, h1.x); a[5] = fmaf(a[5], s1.y, h1.y); a[6] = fmaf(a[6], s1.z, h1.z); a[7] = fmaf(a[7], s1.w, h1.w);
    if (residual) {
      float r[8];
      unpack8(ld_nc_v4(residual + i * 8), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += r[j];
    }
    if (mask) {             // 1 bit per element: the ReLU mask the backward needs (16x smaller than reading z back)
      unsigned m = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) m |= (a[j] > 0.f ? 1u : 0u) << j;
      mask[i] = (unsigned char)m;
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], 0.f);
    }
    if (z) *reinterpret_cast<uint4*>(z + i * 8) = pack8(a);
    if (z_f32) {
      *reinterpret_cast<float4*>(z_f32 + i * 8) = make_float4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<float4*>(z_f32 + i * 8 + 4) = make_float4(a[4], a[5], a[6], a[7]);
    }
  }
}

// dy = k1*g + k2*y + k3 with g = dz * relu_mask (coef from bn_bwd_finalize_kernel);   d_res = g (optional)
__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dz,
                                    const __nv_bfloat16* __restrict__ z, const float* __restrict__ coef,
                                    __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dres, long long total8,
                                    int C8, int relu) {
  const int C = C8 * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % C8) * 8;
    float a[8], g[8];
    unpack8(ld_nc_v4(y + i * 8), a);
    unpack8(ld_nc_v4(dz + i * 8), g);
    if (relu == 1) {
      float zz[8];
      unpack8(ld_nc_v4(z + i * 8), zz);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = zz[j] > 0.f ? g[j] : 0.f;
    } else if (relu == 3) {       // `z` carries the 1-bit-per-element mask written by bn_apply_mask
      const unsigned mb = reinterpret_cast<const unsigned char*>(z)[i];
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = ((mb >> j) & 1u) ? g[j] : 0.f;
    } else if (relu == 2) {       // mask recomputed from y: `z` carries fp32 [2, C] (scale, shift)
      const float* ss = reinterpret_cast<const float*>(z);
      const float4 sa = *reinterpret_cast<const float4*>(ss + c0), sb = *reinterpret_cast<const float4*>(ss + c0 + 4);
      const float4 ha = *reinterpret_cast<const float4*>(ss + C + c0), hb = *reinterpret_cast<const float4*>(ss + C + c0 + 4);
      const float scv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
      const float shv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = fmaf(a[j], scv[j], shv[j]) > 0.f ? g[j] : 0.f;
    }
    if (dres) *reinterpret_cast<uint4*>(dres + i * 8) = pack8(g);
    const float4 k1a = *reinterpret_cast<const float4*>(coef + c0), k1b = *reinterpret_cast<const float4*>(coef + c0 + 4);
    const float4 k2a = *reinterpret_cast<const float4*>(coef + C + c0), k2b = *reinterpret_cast<const float4*>(coef + C + c0 + 4);
    const float4 k3a = *reinterpret_cast<const float4*>(coef + 2 * C + c0), k3b = *reinterpret_cast<const float4*>(coef + 2 * C + c0 + 4);
    float o[8];
    o[0] = fmaf(k1a.x, g[0], fmaf(k2a.x, a[0], k3a.x)); o[1] = fmaf(k1a.y, g[1], fmaf(k2a.y, a[1], k3a.y));
    o[2] = fmaf(k1a.z, g[2], fmaf(k2a.z, a[2], k3a.z)); o[3] = fmaf(k1a.w, g[3], fmaf(k2a.w, a[3], k3a.w));
    o[4] = fmaf(k1b.x, g[4], fmaf(k2b.x, a[4], k3b.x)); o[5] = fmaf(k1b.y, g[5], fmaf(k2b.y, a[5], k3b.y));
    o[6] = fmaf(k1b.z, g[6], fmaf(k2b.z, a[6], k3b.z)); o[7] = fmaf(k1b.w, g[7], fmaf(k2b.w, a[7], k3b.w));
    *reinterpret_cast<uint4*>(dy + i * 8) = pack8(o);
  }
}

// use_global_stats / eval mode: scale & shift straight from the running statistics (freeze.py:17-23)
__global__ void bn_global_affine_kernel(const float* __restrict__ rmean, const float* __restrict__ rvar,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, float* mean,
                                        float* invstd, float* scale, float* shift, float eps, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float m = rmean[c], is = rsqrtf(rvar[c] + eps);
  float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean[c] = m; invstd[c] = is; scale[c] = g * is; shift[c] = b - m * g * is;
}

__global__ void axpy_f32_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaf(a, x[i], y[i]);
}

// sums[0|1][c] = sum_b part[b][0|1][c];  optionally dbeta += sums[0], dgamma += sums[1];  optionally the per-channel
// coefficients of the BN input gradient  dy = k1*g + k2*y + k3  (k1 = gamma*invstd, k2 = -gamma*invstd^2*mean(g*xhat),
// k3 = -k1*mean(g) - k2*mean):  coef [3, C].  Block (32, 32) like bn_finalize_kernel.
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, int nblk, float* __restrict__ sums,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, const float* __restrict__ gamma,
                                       const float* __restrict__ mean, const float* __restrict__ invstd, float inv_count,
                                       float* __restrict__ coef, int C) {
  __shared__ float r0[32][33], r1[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, b = 0.f;
  if (c < C)
    for (int k = threadIdx.y; k < nblk; k += 32) { a += part[(size_t)k * 2 * C + c]; b += part[(size_t)k * 2 * C + C + c]; }
  r0[threadIdx.y][threadIdx.x] = a; r1[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  for (int h = 16; h > 0; h >>= 1) {
    if ((int)threadIdx.y < h) { r0[threadIdx.y][threadIdx.x] += r0[threadIdx.y + h][threadIdx.x]; r1[threadIdx.y][threadIdx.x] += r1[threadIdx.y + h][threadIdx.x]; }
    __syncthreads();
  }
  if (threadIdx.y != 0 || c >= C) return;
  a = r0[0][threadIdx.x]; b = r1[0][threadIdx.x];
  if (sums) { sums[c] = a; sums[C + c] = b; }
  if (dbeta) dbeta[c] += a;
  if (dgamma) dgamma[c] += b;
  if (coef) {
    const float is = invstd[c], ga = gamma ? gamma[c] : 1.f;
    const float k1 = ga * is;
    const float k2 = -ga * is * is * (b * inv_count);
    coef[c] = k1;
    coef[C + c] = k2;
    coef[2 * C + c] = -k1 * (a * inv_count) - k2 * mean[c];
  }
}

static void reduce_cfg(long long P, int C, dim3& grid, dim3& block, int& rows_per_block, int& smem) {
  int cg = C / 8;
  int gx = 1;
  while (gx * 2 <= cg && gx < 64) gx *= 2;       // power of two <= min(cg, 64)
  int gy = 256 / gx;                              // power of two (tree reduction)
  block = dim3(gx, gy);
  int cgb = (cg + gx - 1) / gx;
  long long target_blocks = (long long)num_sms() * 4 / cgb;
  if (target_blocks < 1) target_blocks = 1;
  long long rpb = (P + target_blocks - 1) / target_blocks;
  if (rpb < gy) rpb = gy;
  rows_per_block = (int)rpb;
  grid = dim3((unsigned)((P + rpb - 1) / rpb), cgb);
  smem = gx * gy * 16 * 4;
}

static int ew_blocks(long long total8) {
  long long g = (total8 + 255) / 256;
  long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace pb

using namespace pb;

// number of row blocks (= partials per channel) the reduce kernels produce for a [P, C] tensor
extern "C" int passl_b200_bn_reduce_blocks(long long P, int C) {
  if (P <= 0 || C <= 0 || C % 8) return 0;
  dim3 grid, block; int rpb, smem;
  reduce_cfg(P, C, grid, block, rpb, smem);
  return (int)grid.x;
}

// part[b][0][c] = sum over the rows of block b of y[p,c];  part[b][1][c] = same for y^2.   part: fp32 [nblk, 2, C]
extern "C" int passl_b200_bn_stats(const void* y, float* part, long long P, int C, void* stream) {
  if (P <= 0 || C <= 0 || C % 8) return PB_ERR_BAD_ARG;
  dim3 grid, block; int rpb, smem;
  reduce_cfg(P, C, grid, block, rpb, smem);
  bn_reduce_kernel<0><<<grid, block, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(y), nullptr, nullptr,
                                                                    nullptr, nullptr, part, P, C, rpb, 0);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_bn_finalize(const float* part, int nblk, const float* gamma, const float* beta,
                                      float* mean, float* invstd, float* scale, float* shift, float* running_mean,
                                      float* running_var, long long count, float eps, float momentum, int C, void* stream) {
  if (C <= 0 || count <= 0 || nblk <= 0) return PB_ERR_BAD_ARG;
  bn_finalize_kernel<<<(C + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream>>>(part, nblk, gamma, beta, mean, invstd, scale, shift,
                                                                        running_mean, running_var, 1.f / (float)count, eps,
                                                                        momentum, C);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_bn_global_affine(const float* running_mean, const float* running_var, const float* gamma,
                                           const float* beta, float* mean, float* invstd, float* scale, float* shift,
                                           float eps, int C, void* stream) {
  if (C <= 0) return PB_ERR_BAD_ARG;
  bn_global_affine_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, gamma, beta, mean,
                                                                             invstd, scale, shift, eps, C);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

// y += a * x   (fp32; small vectors: bias / BN parameter gradients, loss bookkeeping)
extern "C" int passl_b200_axpy_f32(float* y, const float* x, float a, long long n, void* stream) {
  if (n <= 0) return PB_OK;
  long long g = (n + 255) / 256;
  if (g > 1184) g = 1184;
  axpy_f32_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(y, x, a, n);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_bn_apply(const void* y, const void* residual, const float* scale, const float* shift, void* z,
                                   float* z_f32, long long P, int C, int relu, void* stream) {
  if (P <= 0 || C <= 0 || C % 8) return PB_ERR_BAD_ARG;
  long long total8 = P * (C / 8);
  bn_apply_kernel<<<ew_blocks(total8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(y), reinterpret_cast<const __nv_bfloat16*>(residual), scale, shift,
      reinterpret_cast<__nv_bfloat16*>(z), z_f32, total8, C / 8, relu, nullptr);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

// same + relu_mask [P*C/8] bytes: bit j of byte i = (pre-ReLU value of element 8i+j > 0)
extern "C" int passl_b200_bn_apply_mask(const void* y, const void* residual, const float* scale, const float* shift, void* z,
                                        void* relu_mask, long long P, int C, int relu, void* stream) {
  if (P <= 0 || C <= 0 || C % 8 || !relu_mask) return PB_ERR_BAD_ARG;
  long long total8 = P * (C / 8);
  bn_apply_kernel<<<ew_blocks(total8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(y), reinterpret_cast<const __nv_bfloat16*>(residual), scale, shift,
      reinterpret_cast<__nv_bfloat16*>(z), nullptr, total8, C / 8, relu, reinterpret_cast<unsigned char*>(relu_mask));
  PB_LAUNCH_CHECK();
  return PB_OK;
}

// part[b][0][c] = sum_p g[p,c];  part[b][1][c] = sum_p g[p,c]*xhat[p,c]   with g = dz * (z > 0 if relu);  part [nblk,2,C]
extern "C" int passl_b200_bn_bwd_reduce(const void* y, const void* dz, const void* z, const float* mean, const float* invstd,
                                        float* part, long long P, int C, int relu, void* stream) {
  if (P <= 0 || C <= 0 || C % 8) return PB_ERR_BAD_ARG;
  if (relu && !z) return PB_ERR_BAD_ARG;
  dim3 grid, block; int rpb, smem;
  reduce_cfg(P, C, grid, block, rpb, smem);
  if (relu == 3)
    bn_reduce_kernel<3><<<grid, block, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(y), reinterpret_cast<const __nv_bfloat16*>(dz),
        reinterpret_cast<const __nv_bfloat16*>(z), mean, invstd, part, P, C, rpb, relu);
  else if (relu == 2)
    bn_reduce_kernel<2><<<grid, block, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(y), reinterpret_cast<const __nv_bfloat16*>(dz),
        reinterpret_cast<const __nv_bfloat16*>(z), mean, invstd, part, P, C, rpb, relu);
  else
    bn_reduce_kernel<1><<<grid, block, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(y), reinterpret_cast<const __nv_bfloat16*>(dz),
        reinterpret_cast<const __nv_bfloat16*>(z), mean, invstd, part, P, C, rpb, relu);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

// sums [2, C] = totals of the partials; dbeta += sums[0], dgamma += sums[1] when given (fp32 gradient accumulators)
extern "C" int passl_b200_bn_bwd_finalize(const float* part, int nblk, float* sums, float* dgamma, float* dbeta,
                                          const float* gamma, const float* mean, const float* invstd, long long count,
                                          float* coef, int C, void* stream) {
  if (C <= 0 || nblk <= 0) return PB_ERR_BAD_ARG;
  bn_bwd_finalize_kernel<<<(C + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream>>>(part, nblk, sums, dgamma, dbeta, gamma, mean,
                                                                                    invstd, count > 0 ? 1.f / (float)count : 0.f,
                                                                                    coef, C);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_bn_bwd_apply(const void* y, const void* dz, const void* z, const float* coef, void* dy, void* dres,
                                       long long P, int C, int relu, void* stream) {
  if (P <= 0 || C <= 0 || C % 8 || !coef) return PB_ERR_BAD_ARG;
  if (relu && !z) return PB_ERR_BAD_ARG;
  long long total8 = P * (C / 8);
  bn_bwd_apply_kernel<<<ew_blocks(total8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(y), reinterpret_cast<const __nv_bfloat16*>(dz),
      reinterpret_cast<const __nv_bfloat16*>(z), coef, reinterpret_cast<__nv_bfloat16*>(dy),
      reinterpret_cast<__nv_bfloat16*>(dres), total8, C / 8, relu);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
