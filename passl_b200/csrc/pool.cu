// Pooling and input-layout kernels of the ResNet / ViT stems (channels-last bf16).
//   * im2col from the reference's NCHW fp32 image batch into a bf16 [pixels, Kpad] matrix with (r, s, c) K order —
//     feeds the tcgen05 GEMM for the 7x7/2 stem conv (resnetimagenet.py:190-198, K=147 -> 152) and the 16x16/16 ViT
//     patch embedding (passl/models/vision_transformer.py:231-236).  Fuses layout change + cast.
//   * 3x3/2 max-pool forward (saves the arg-max tap as int8) and backward (gather form, no atomics).
//     Tie rule = first maximum in (h, w) scan order, like Paddle's MaxPool2dGradFunctor.
//   * global average pool forward / backward (AdaptiveAvgPool2D((1,1)): necks/base_neck.py:52,79).
#include "common.cuh"
#include "host_utils.h"
#include "../../include/passl_b200.h"

namespace pb {
// (i -> cg, w, h, n) of a [N, H, W, C/8] index.  64-bit div / mod are ~100-instruction emulation sequences; three of them per
// 16-byte output made the pooling kernels instruction-bound (1.1 / 1.4 ms for tensors that stream in 0.37 ms).
__device__ __forceinline__ void decode_nhwc8(long long i, int C8, int W, int H, int& cg, int& w, int& h, int& n) {
  if (i < 0x7fffffffLL) {
    unsigned u = (unsigned)i;
    cg = (int)(u % (unsigned)C8); u /= (unsigned)C8;
    w = (int)(u % (unsigned)W); u /= (unsigned)W;
    h = (int)(u % (unsigned)H);
    n = (int)(u / (unsigned)H);
  } else {
    cg = (int)(i % C8);
    long long p = i / C8;
    w = (int)(p % W);
    p /= W;
    h = (int)(p % H);
    n = (int)(p / H);
  }
}


__device__ __forceinline__ void unpack8p(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8p(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// out[(n, oh, ow), (r*S + s)*C + c] = x[n, c, oh*stride + r - pad, ow*stride + s - pad]   (0 outside / in K padding)
__global__ void im2col_nchw_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C, int H, int W,
                                   int R, int S, int stride, int pad, int Ho, int Wo, int Kpad) {
  const int chunks = Kpad / 8;
  const long long total = (long long)N * Ho * Wo * chunks;
  const int Kreal = R * S * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    long long p = i / chunks;
    const int ow = (int)(p % Wo);
    p /= Wo;
    const int oh = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = ch * 8 + j;
      float val = 0.f;
      if (k < Kreal) {
        const int c = k % C;
        const int rs = k / C;
        const int s = rs % S, r = rs / S;
        const int h = oh * stride + r - pad, w = ow * stride + s - pad;
        if (h >= 0 && h < H && w >= 0 && w < W) val = __ldg(x + (((long long)n * C + c) * H + h) * W + w);
      }
      v[j] = val;
    }
    *reinterpret_cast<uint4*>(out + i * 8) = pack8p(v);
  }
}

// NHWC 3x3 stride 2 pad 1 max pool; 8 channels / thread
__global__ void maxpool3x3s2_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                        signed char* __restrict__ argmax, int N, int H, int W, int C, int Ho, int Wo) {
  const int C8 = C / 8;
  const long long total = (long long)N * Ho * Wo * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cg, ow, oh, n;
    decode_nhwc8(i, C8, Wo, Ho, cg, ow, oh, n);
    float best[8];
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = -1; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = oh * 2 + r - 1;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = ow * 2 + s - 1;
        if (w < 0 || w >= W) continue;
        float v[8];
        // L1-allocating load: the 3x3/2 windows of neighbouring outputs overlap (each input pixel is read 2.25x)
        unpack8p(__ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * C + cg * 8)), v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[j] > best[j] || arg[j] < 0) { best[j] = v[j]; arg[j] = r * 3 + s; }
      }
    }
    *reinterpret_cast<uint4*>(y + i * 8) = pack8p(best);
    uint2 a;
    a.x = (arg[0] & 0xff) | ((arg[1] & 0xff) << 8) | ((arg[2] & 0xff) << 16) | ((arg[3] & 0xff) << 24);
    a.y = (arg[4] & 0xff) | ((arg[5] & 0xff) << 8) | ((arg[6] & 0xff) << 16) | ((arg[7] & 0xff) << 24);
    *reinterpret_cast<uint2*>(argmax + i * 8) = a;
  }
}

// Stem tail fused: out = maxpool3x3/2(relu(y * scale + shift)) without materialising the normalised activation (the BN backward
// recomputes the ReLU mask from y, the pool backward only needs the arg-max tap): saves one write + one read of [N,112,112,64].
__global__ void bn_relu_maxpool3x3s2_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                                                const float* __restrict__ shift, __nv_bfloat16* __restrict__ y,
                                                signed char* __restrict__ argmax, int N, int H, int W, int C, int Ho, int Wo) {
  const int C8 = C / 8;
  const long long total = (long long)N * Ho * Wo * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cg, ow, oh, n;
    decode_nhwc8(i, C8, Wo, Ho, cg, ow, oh, n);
    float sc[8], sh[8];
#pragma unroll
    for (int j4 = 0; j4 < 2; ++j4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(scale + cg * 8) + j4), b = __ldg(reinterpret_cast<const float4*>(shift + cg * 8) + j4);
      sc[j4 * 4 + 0] = a.x; sc[j4 * 4 + 1] = a.y; sc[j4 * 4 + 2] = a.z; sc[j4 * 4 + 3] = a.w;
      sh[j4 * 4 + 0] = b.x; sh[j4 * 4 + 1] = b.y; sh[j4 * 4 + 2] = b.z; sh[j4 * 4 + 3] = b.w;
    }
    float best[8];
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = -1; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = oh * 2 + r - 1;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = ow * 2 + s - 1;
        if (w < 0 || w >= W) continue;
        float v[8];
        unpack8p(__ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * C + cg * 8)), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // same rounding as bn_apply followed by the pool: the pooled tensor holds bf16(relu(fma))
          const float z = __bfloat162float(__float2bfloat16_rn(fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f)));
          if (z > best[j] || arg[j] < 0) { best[j] = z; arg[j] = r * 3 + s; }
        }
      }
    }
    *reinterpret_cast<uint4*>(y + i * 8) = pack8p(best);
    uint2 a;
    a.x = (arg[0] & 0xff) | ((arg[1] & 0xff) << 8) | ((arg[2] & 0xff) << 16) | ((arg[3] & 0xff) << 24);
    a.y = (arg[4] & 0xff) | ((arg[5] & 0xff) << 8) | ((arg[6] & 0xff) << 16) | ((arg[7] & 0xff) << 24);
    *reinterpret_cast<uint2*>(argmax + i * 8) = a;
  }
}

__global__ void maxpool3x3s2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const signed char* __restrict__ argmax,
                                        __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo) {
  const int C8 = C / 8;
  const long long total = (long long)N * H * W * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cg, w, h, n;
    decode_nhwc8(i, C8, W, H, cg, w, h, n);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int oh_lo = (h) / 2 + ((h & 1) ? 0 : 0);  // windows oh with oh*2-1 <= h <= oh*2+1
    for (int oh = (h - 1 + 1) / 2; oh <= (h + 1) / 2; ++oh) {
      if (oh < 0 || oh >= Ho) continue;
      const int r = h - (oh * 2 - 1);
      if (r < 0 || r > 2) continue;
      for (int ow = (w) / 2; ow <= (w + 1) / 2; ++ow) {
        if (ow < 0 || ow >= Wo) continue;
        const int s = w - (ow * 2 - 1);
        if (s < 0 || s > 2) continue;
        const long long o = (((long long)n * Ho + oh) * Wo + ow) * C8 + cg;
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(argmax + o * 8));   // cached: 4 input pixels share each window
        float g[8];
        unpack8p(__ldg(reinterpret_cast<const uint4*>(dy + o * 8)), g);
        const int tap = r * 3 + s;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int aj = (j < 4 ? (a.x >> (8 * j)) : (a.y >> (8 * (j - 4)))) & 0xff;
          if (aj == tap) acc[j] += g[j];
        }
      }
    }
    (void)oh_lo;
    *reinterpret_cast<uint4*>(dx + i * 8) = pack8p(acc);
  }
}

// y[n, c] = mean_{hw} x[n, hw, c];  one thread per (n, 8 channels)
__global__ void avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y_bf16,
                                   float* __restrict__ y_f32, int N, int HW, int C) {
  const int C8 = C / 8;
  const int total = N * C8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cg = i % C8, n = i / C8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int p = 0; p < HW; ++p) {
      float v[8];
      unpack8p(ld_nc_v4(x + ((long long)n * HW + p) * C + cg * 8), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    const float inv = 1.f / HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    if (y_bf16) *reinterpret_cast<uint4*>(y_bf16 + (long long)i * 8) = pack8p(acc);
    if (y_f32) {
      *reinterpret_cast<float4*>(y_f32 + (long long)i * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(y_f32 + (long long)i * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}
// dx[n, hw, c] = dy[n, c] / HW
__global__ void avgpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N, int HW, int C) {
  const int C8 = C / 8;
  const long long total = (long long)N * HW * C8;
  const float inv = 1.f / HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C8);
    const long long n = i / ((long long)HW * C8);
    float g[8];
    unpack8p(*reinterpret_cast<const uint4*>(dy + (n * C8 + cg) * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= inv;
    *reinterpret_cast<uint4*>(dx + i * 8) = pack8p(g);
  }
}

static int ew_blocks_p(long long n) {
  long long g = (n + 255) / 256;
  long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace pb

using namespace pb;

extern "C" int passl_b200_im2col_nchw_f32(const float* x, void* out, int N, int C, int H, int W, int R, int S, int stride,
                                          int pad, int Kpad, void* stream) {
  if (N <= 0 || Kpad % 8 || Kpad < R * S * C) return PB_ERR_BAD_ARG;
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
  long long total = (long long)N * Ho * Wo * (Kpad / 8);
  im2col_nchw_kernel<<<ew_blocks_p(total), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), N, C, H,
                                                                           W, R, S, stride, pad, Ho, Wo, Kpad);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
extern "C" int passl_b200_maxpool3x3s2_fwd(const void* x, void* y, void* argmax, int N, int H, int W, int C, void* stream) {
  if (N <= 0 || C % 8) return PB_ERR_BAD_ARG;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * Ho * Wo * (C / 8);
  maxpool3x3s2_fwd_kernel<<<ew_blocks_p(total), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), reinterpret_cast<signed char*>(argmax), N,
      H, W, C, Ho, Wo);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
extern "C" int passl_b200_bn_relu_maxpool3x3s2_fwd(const void* x, const float* scale, const float* shift, void* y, void* argmax, int N,
                                                  int H, int W, int C, void* stream) {
  if (N <= 0 || C % 8 || !scale || !shift) return PB_ERR_BAD_ARG;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * Ho * Wo * (C / 8);
  bn_relu_maxpool3x3s2_fwd_kernel<<<ew_blocks_p(total), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), scale, shift, reinterpret_cast<__nv_bfloat16*>(y),
      reinterpret_cast<signed char*>(argmax), N, H, W, C, Ho, Wo);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
extern "C" int passl_b200_maxpool3x3s2_bwd(const void* dy, const void* argmax, void* dx, int N, int H, int W, int C,
                                           void* stream) {
  if (N <= 0 || C % 8) return PB_ERR_BAD_ARG;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * H * W * (C / 8);
  maxpool3x3s2_bwd_kernel<<<ew_blocks_p(total), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const signed char*>(argmax),
      reinterpret_cast<__nv_bfloat16*>(dx), N, H, W, C, Ho, Wo);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
extern "C" int passl_b200_avgpool_fwd(const void* x, void* y_bf16, float* y_f32, int N, int HW, int C, void* stream) {
  if (N <= 0 || C % 8 || HW <= 0) return PB_ERR_BAD_ARG;
  avgpool_fwd_kernel<<<ew_blocks_p((long long)N * C / 8), 128, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, N, HW, C);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
extern "C" int passl_b200_avgpool_bwd(const void* dy, void* dx, int N, int HW, int C, void* stream) {
  if (N <= 0 || C % 8 || HW <= 0) return PB_ERR_BAD_ARG;
  avgpool_bwd_kernel<<<ew_blocks_p((long long)N * HW * C / 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<__nv_bfloat16*>(dx), N, HW, C);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
