// ResNet stem 7x7/2 (3 -> 64, pad 3) without an im2col matrix (resnetimagenet.py:190-198, `self.conv1`).
//
// The reference hands NCHW fp32 images to cuDNN.  A 3-channel input cannot feed TMA / tcgen05 directly (6 B pixels), and a full
// im2col of 224^2 images is 3.9 GB per 1024 images.  Instead the image is repacked once into a "W-unfolded space-to-depth" tensor
//     xp[n, i, q, ((dj*2 + a)*2 + b)*4 + c] = img[n, c, 2i + a, 2(q + dj - 2) + b]        (0 outside, c == 3 is padding)
// [N, 112, 112, 64] bf16 (1.64 GB / 1024 images), on which the stem is an ordinary 4x1 implicit-GEMM convolution with 64 input
// channels (rows padded (2, 1)):  out[n, p, q, co] = sum_{di, ch} xp[n, p + di - 2, q, ch] * wp[co, di, ch],
//     wp[co, di, ((dj*2 + a)*2 + b)*4 + c] = w[co, r = 2di + a - 1, s = 2dj + b - 1, c]      (0 when r < 0, s < 0 or c == 3)
// because 2p + r - 3 = 2(p + di - 2) + a.  The weight-gradient GEMM reads the same xp; its result is folded back into the
// [64, (r, s, c)] gradient by stem_unpack_wgrad.
#include "common.cuh"
#include "host_utils.h"
#include "../../include/passl_b200.h"

namespace pb {

__global__ void stem_pack_input_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ xp, int N, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * 8;               // one thread per (pixel, 8-channel group = (dj, a))
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t & 7);
    long long pix = t >> 3;
    const int q = (int)(pix % Wo);
    pix /= Wo;
    const int i = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    const int dj = g >> 1, a = g & 1;
    const int row = 2 * i + a;
    const int col0 = 2 * (q + dj - 2);
    float v[8];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int col = col0 + b;
      const bool ok = col >= 0 && col < W;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[b * 4 + c] = ok ? __ldg(img + (((long long)n * 3 + c) * H + row) * W + col) : 0.f;
      v[b * 4 + 3] = 0.f;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(xp + t * 8) = o;
  }
}

// w fp32 [64, kpad] with (r, s, c) order -> wp bf16 [64, 4, 64]
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int kpad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * 256) return;
  const int ch = idx & 63, di = (idx >> 6) & 3, co = idx >> 8;
  const int c = ch & 3, b = (ch >> 2) & 1, a = (ch >> 3) & 1, dj = ch >> 4;
  const int r = 2 * di + a - 1, s = 2 * dj + b - 1;
  float v = 0.f;
  if (c < 3 && r >= 0 && s >= 0) v = w[(size_t)co * kpad + (r * 7 + s) * 3 + c];
  wp[idx] = __float2bfloat16_rn(v);
}

// dw[co, (r, s, c)] += dwp[co, di, ch]
__global__ void stem_unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int kpad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * 147) return;
  const int k = idx % 147, co = idx / 147;
  const int c = k % 3, rs = k / 3, s = rs % 7, r = rs / 7;
  const int di = (r + 1) >> 1, a = (r + 1) & 1, dj = (s + 1) >> 1, b = (s + 1) & 1;
  const int ch = ((dj * 2 + a) * 2 + b) * 4 + c;
  dw[(size_t)co * kpad + k] += dwp[(size_t)co * 256 + di * 64 + ch];
}

}  // namespace pb

using namespace pb;

extern "C" int passl_b200_stem_pack_input(const float* img, void* xp, int N, int H, int W, void* stream) {
  if (N <= 0 || (H & 1) || (W & 1)) return PB_ERR_BAD_ARG;
  const long long total = (long long)N * (H / 2) * (W / 2) * 8;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  stem_pack_input_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(img, reinterpret_cast<__nv_bfloat16*>(xp), N, H, W);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_stem_pack_weight(const float* w, void* wp, int kpad, void* stream) {
  if (kpad < 147) return PB_ERR_BAD_ARG;
  stem_pack_weight_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<__nv_bfloat16*>(wp), kpad);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_stem_unpack_wgrad(const float* dwp, float* dw, int kpad, void* stream) {
  if (kpad < 147) return PB_ERR_BAD_ARG;
  stem_unpack_wgrad_kernel<<<(64 * 147 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dwp, dw, kpad);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
