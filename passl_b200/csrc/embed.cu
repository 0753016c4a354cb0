// Embedding-side kernels of the contrastive heads: row L2-normalisation (fwd/bwd), the MoCo key queue as an
// on-device ring buffer, and the momentum (EMA) key-encoder update over a flat parameter buffer.
//
// Reference: passl/nn/norm.py:18-40 (l2_normalize, sum-form eps), paddle F.normalize (max-form eps, used at
// passl_v110/modeling/architectures/moco.py:159,170), moco.py:77-105 (queue init / _dequeue_and_enqueue),
// moco.py:82-90 (_momentum_update_key_encoder).
#include "common.cuh"
#include "host_utils.h"
#include "../../include/passl_b200.h"

namespace pb {

// one warp per row; D <= 4096
// mode 0: y = x / max(||x||, eps)          (paddle.nn.functional.normalize)
// mode 1: y = x / sqrt(sum x^2 + eps)      (passl.nn.norm.l2_normalize / fluid.layers.l2_normalize)
// mode 2: y = x / ||x||                    (CLIP: clip.py:325-328, no eps)
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, __nv_bfloat16* __restrict__ y_bf16,
                                  float* __restrict__ inv_norm, int N, int D, int mode, float eps) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* xr = x + (size_t)row * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s = fmaf(xr[d], xr[d], s);
  s = warp_sum(s);
  float inv;
  if (mode == 0) inv = 1.f / fmaxf(sqrtf(s), eps);
  else if (mode == 1) inv = 1.f / sqrtf(s + eps);
  else inv = 1.f / sqrtf(s);
  for (int d = lane; d < D; d += 32) {
    float v = xr[d] * inv;
    if (y) y[(size_t)row * D + d] = v;
    if (y_bf16) y_bf16[(size_t)row * D + d] = __float2bfloat16_rn(v);
  }
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
}

// dx = inv * (dy - y * <dy, y> * c)   with c = 1 for modes 1/2 (and mode 0 when ||x|| > eps);
// mode 1: y = x * (s+eps)^-1/2  ->  dx = inv*dy - x * inv^3 * <dy, x> = inv * (dy - y <dy,y>)   (exact)
// mode 0 with ||x|| <= eps: y = x/eps -> dx = dy/eps.
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                  const float* __restrict__ inv_norm, float* __restrict__ dx,
                                  __nv_bfloat16* __restrict__ dx_bf16, int N, int D, int mode, float eps) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* yr = y + (size_t)row * D;
  const float* gr = dy + (size_t)row * D;
  float inv = inv_norm[row];
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) dot = fmaf(gr[d], yr[d], dot);
  dot = warp_sum(dot);
  if (mode == 0 && inv >= 1.f / eps) dot = 0.f;  // clamped branch: pure scaling
  for (int d = lane; d < D; d += 32) {
    float v = inv * (gr[d] - yr[d] * dot);
    if (dx) dx[(size_t)row * D + d] = v;
    if (dx_bf16) dx_bf16[(size_t)row * D + d] = __float2bfloat16_rn(v);
  }
}

// Ring-buffer enqueue. queue is stored key-major [K, D] (the reference keeps [D, K]; converted at checkpoint I/O).
//   queue[ptr + i, :] = keys[i, :]  for i < Bg ;  ptr <- (ptr + Bg) % K      (ptr is int64 on device: no host sync)
// The pointer update is done by a second tiny kernel so every CTA of the copy sees the old value.
__global__ void queue_enqueue_kernel(const float* __restrict__ keys, float* __restrict__ q_f32,
                                     __nv_bfloat16* __restrict__ q_bf16, const long long* __restrict__ ptr, int Bg,
                                     int D, int K) {
  const long long p0 = ptr[0];
  const long long total = (long long)Bg * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / D, d = i - r * D;
    long long dst = ((p0 + r) % K) * D + d;
    float v = keys[i];
    if (q_f32) q_f32[dst] = v;
    if (q_bf16) q_bf16[dst] = __float2bfloat16_rn(v);
  }
}
__global__ void queue_ptr_advance_kernel(long long* ptr, int Bg, int K) { ptr[0] = (ptr[0] + Bg) % K; }

// k = m*k + (1-m)*q over a flat fp32 buffer; optionally refresh the bf16 compute copy of k.
__global__ void ema_update_kernel(float* __restrict__ k, const float* __restrict__ q, __nv_bfloat16* __restrict__ k_bf16,
                                  float m, long long n) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 a = *reinterpret_cast<float4*>(k + i);
    float4 b = *reinterpret_cast<const float4*>(q + i);
    a.x = a.x * m + b.x * (1.f - m);
    a.y = a.y * m + b.y * (1.f - m);
    a.z = a.z * m + b.z * (1.f - m);
    a.w = a.w * m + b.w * (1.f - m);
    *reinterpret_cast<float4*>(k + i) = a;
    if (k_bf16) {
      uint2 u;
      u.x = pack_bf16x2(a.x, a.y);
      u.y = pack_bf16x2(a.z, a.w);
      *reinterpret_cast<uint2*>(k_bf16 + i) = u;
    }
  }
  // tail (n not multiple of 4): handled by the first thread
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long j = n & ~3LL; j < n; ++j) {
      float v = k[j] * m + q[j] * (1.f - m);
      k[j] = v;
      if (k_bf16) k_bf16[j] = __float2bfloat16_rn(v);
    }
  }
}

__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 a = *reinterpret_cast<const float4*>(x + i);
    uint2 u;
    u.x = pack_bf16x2(a.x, a.y);
    u.y = pack_bf16x2(a.z, a.w);
    *reinterpret_cast<uint2*>(y + i) = u;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long j = n & ~3LL; j < n; ++j) y[j] = __float2bfloat16_rn(x[j]);
}
__global__ void cast_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = __bfloat162float(x[i]);
}

static int ew_grid(long long n, int per_thread, int block) {
  long long g = (n / per_thread + block - 1) / block;
  long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace pb

using namespace pb;

extern "C" int passl_b200_l2norm_fwd(const float* x, float* y, void* y_bf16, float* inv_norm, int N, int D, int mode,
                                     float eps, void* stream) {
  if (N <= 0 || D <= 0 || mode < 0 || mode > 2) return PB_ERR_BAD_ARG;
  l2norm_fwd_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, y, reinterpret_cast<__nv_bfloat16*>(y_bf16),
                                                                   inv_norm, N, D, mode, eps);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
extern "C" int passl_b200_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, void* dx_bf16,
                                     int N, int D, int mode, float eps, void* stream) {
  if (N <= 0 || D <= 0 || mode < 0 || mode > 2) return PB_ERR_BAD_ARG;
  l2norm_bwd_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(dy, y, inv_norm, dx,
                                                                   reinterpret_cast<__nv_bfloat16*>(dx_bf16), N, D, mode, eps);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
extern "C" int passl_b200_queue_enqueue(const float* keys, float* queue_f32, void* queue_bf16, long long* queue_ptr,
                                        int Bg, int D, int K, void* stream) {
  if (Bg <= 0 || D <= 0 || K <= 0) return PB_ERR_BAD_ARG;
  if (K % Bg != 0) return PB_ERR_BAD_ARG;  // moco.py:99  assert self.K % batch_size == 0
  cudaStream_t st = (cudaStream_t)stream;
  queue_enqueue_kernel<<<ew_grid((long long)Bg * D, 1, 256), 256, 0, st>>>(
      keys, queue_f32, reinterpret_cast<__nv_bfloat16*>(queue_bf16), queue_ptr, Bg, D, K);
  PB_LAUNCH_CHECK();
  queue_ptr_advance_kernel<<<1, 1, 0, st>>>(queue_ptr, Bg, K);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
extern "C" int passl_b200_ema_update(float* k, const float* q, void* k_bf16, float m, long long n, void* stream) {
  if (n <= 0) return PB_OK;
  if ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(q)) & 15) return PB_ERR_BAD_ARG;
  ema_update_kernel<<<ew_grid(n, 4, 256), 256, 0, (cudaStream_t)stream>>>(k, q, reinterpret_cast<__nv_bfloat16*>(k_bf16), m, n);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
extern "C" int passl_b200_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream) {
  if (n <= 0) return PB_OK;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 7)) return PB_ERR_BAD_ARG;
  cast_f32_to_bf16_kernel<<<ew_grid(n, 4, 256), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
extern "C" int passl_b200_cast_bf16_to_f32(const void* x, float* y, long long n, void* stream) {
  if (n <= 0) return PB_OK;
  cast_bf16_to_f32_kernel<<<ew_grid(n, 1, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), y, n);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
