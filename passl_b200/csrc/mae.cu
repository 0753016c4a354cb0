// MAE-specific kernels (passl/models/mae.py): per-sample random masking, token assembly (gather + positional embedding +
// cls / mask tokens) with its backward, and the fused masked-patch MSE loss.
//
//   random_masking   mae.py:184-212  noise -> argsort -> ids_shuffle / ids_restore (int64, bit-exact for a given noise), mask
//   token assembly   mae.py:214-266  "x + pos_embed[:,1:]" -> gather kept tokens -> prepend cls (encoder);
//                                    concat mask tokens -> un-shuffle by ids_restore -> prepend cls -> + decoder_pos_embed
//   forward_loss     mae.py:268-284  patchify('nchpwq->nhwpqc') -> optional per-patch (t-mean)/sqrt(var+1e-6) (unbiased var)
//                                    -> mean((pred-t)^2, -1) -> sum(loss*mask)/sum(mask)
// HBM-bound; the loss reads only the masked patches (0.75*B*196*768*(s_pred + s_img) bytes).
#include "common.cuh"
#include "host_utils.h"
#include "../../include/passl_b200.h"

namespace pb {

// ---------------------------------------------------------------------------------------------------------------------
// random masking: one CTA per sample, bitonic sort of (noise, index) in shared memory (L <= 1024)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void mae_masking_kernel(const float* __restrict__ noise, long long* __restrict__ ids_shuffle,
                                   long long* __restrict__ ids_restore, float* __restrict__ mask, int L, int len_keep, int LP) {
  extern __shared__ unsigned char sm_raw[];
  float* key = reinterpret_cast<float*>(sm_raw);
  int* idx = reinterpret_cast<int*>(key + LP);
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < LP; i += blockDim.x) {
    key[i] = i < L ? noise[(size_t)b * L + i] : INFINITY;
    idx[i] = i;
  }
  __syncthreads();
  for (int k = 2; k <= LP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < LP; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = ((i & k) == 0);
          const float a = key[i], c = key[ixj];
          const int ia = idx[i], ic = idx[ixj];
          const bool gt = (a > c) || (a == c && ia > ic);   // stable: ties broken by original index
          if (gt == up) { key[i] = c; key[ixj] = a; idx[i] = ic; idx[ixj] = ia; }
        }
      }
      __syncthreads();
    }
  }
  for (int r = threadIdx.x; r < L; r += blockDim.x) {
    const int s = idx[r];                       // ids_shuffle[r] = s : the r-th smallest noise sits at position s
    ids_shuffle[(size_t)b * L + r] = s;
    ids_restore[(size_t)b * L + s] = r;
    mask[(size_t)b * L + s] = r < len_keep ? 0.f : 1.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// token assembly.  mode 0 (ViT):      out[b,0] = cls + pos[0];  out[b,t] = src[b,t-1] + pos[t]
//                  mode 1 (MAE enc):  out[b,0] = cls + pos[0];  out[b,1+j] = src[b,s] + pos[1+s],  s = ids_shuffle[b,j], j < keep
//                  mode 2 (MAE dec):  out[b,0] = src[b,0] + pos[0];  out[b,1+i] = (r < keep ? src[b,1+r] : mask_token) + pos[1+i],
//                                     r = ids_restore[b,i]
// src bf16 [B, Ls, D]; out bf16 [B, Lo, D]; pos fp32 [>= Lo or L+1, D]; tok fp32 [D] (cls or mask token).  8 channels/thread.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld8f(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void unpack8m(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8m(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

__global__ void token_assemble_fwd_kernel(const __nv_bfloat16* __restrict__ src, const long long* __restrict__ ids,
                                          const float* __restrict__ pos, const float* __restrict__ tok,
                                          __nv_bfloat16* __restrict__ out, int B, int Ls, int Lo, int D, int mode, int keep) {
  const int D8 = D / 8;
  const long long total = (long long)B * Lo * D8;
  const int L = (mode == 1) ? Ls : Lo - 1;      // ids row length (number of patches)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % D8) * 8;
    const long long bt = i / D8;
    const int t = (int)(bt % Lo), b = (int)(bt / Lo);
    float v[8], pe[8];
    int prow = t;
    const __nv_bfloat16* sp = nullptr;
    if (mode == 0) {
      if (t > 0) sp = src + ((size_t)b * Ls + (t - 1)) * D;
    } else if (mode == 1) {
      if (t > 0) { const int s = (int)ids[(size_t)b * L + (t - 1)]; sp = src + ((size_t)b * Ls + s) * D; prow = 1 + s; }
    } else {
      if (t == 0) sp = src + (size_t)b * Ls * D;
      else { const int r = (int)ids[(size_t)b * L + (t - 1)]; if (r < keep) sp = src + ((size_t)b * Ls + 1 + r) * D; }
    }
    if (sp) unpack8m(*reinterpret_cast<const uint4*>(sp + c0), v);
    else ld8f(tok + c0, v);
    ld8f(pos + (size_t)prow * D + c0, pe);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += pe[j];
    *reinterpret_cast<uint4*>(out + bt * D + c0) = pack8m(v);
  }
}

// backward in gather form (every source row is used at most once -> no atomics, no zero-init):
//   mode 0: dsrc[b,s] = dout[b,1+s]
//   mode 1: dsrc[b,s] = (r = ids_restore[b,s]) < keep ? dout[b,1+r] : 0
//   mode 2: dsrc[b,0] = dout[b,0];  dsrc[b,1+r] = dout[b, 1+ids_shuffle[b,r]]   (r < keep)
__global__ void token_assemble_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const long long* __restrict__ ids,
                                          __nv_bfloat16* __restrict__ dsrc, int B, int Ls, int Lo, int D, int mode, int keep) {
  const int D8 = D / 8;
  const long long total = (long long)B * Ls * D8;
  const int L = (mode == 1) ? Ls : Lo - 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % D8) * 8;
    const long long bs = i / D8;
    const int s = (int)(bs % Ls), b = (int)(bs / Ls);
    int t = -1;
    if (mode == 0) t = 1 + s;
    else if (mode == 1) { const int r = (int)ids[(size_t)b * L + s]; t = r < keep ? 1 + r : -1; }
    else t = (s == 0) ? 0 : 1 + (int)ids[(size_t)b * L + (s - 1)];
    uint4 u = make_uint4(0, 0, 0, 0);
    if (t >= 0) u = *reinterpret_cast<const uint4*>(dout + ((size_t)b * Lo + t) * D + c0);
    *reinterpret_cast<uint4*>(dsrc + bs * D + c0) = u;
  }
}

// gradients of the broadcast token (cls: rows t == 0; mask token (mode 2): rows with ids_restore >= keep) and, optionally, of a
// learnable positional table: acc_tok[c] += sum, acc_pos[t, c] += sum_b dout[b, t, c].  One CTA per 8-channel group.
__global__ void __launch_bounds__(256) token_param_grad_kernel(const __nv_bfloat16* __restrict__ dout, const long long* __restrict__ ids,
                                                               float* __restrict__ acc_tok, float* __restrict__ acc_pos, int B,
                                                               int Lo, int D, int mode, int keep) {
  __shared__ float red[256][8];
  const int c0 = blockIdx.x * 8;
  const int L = Lo - 1;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (acc_tok) {
    if (mode == 2) {
      for (long long r = threadIdx.x; r < (long long)B * L; r += blockDim.x) {
        const int b = (int)(r / L), i = (int)(r % L);
        if ((int)ids[r] >= keep) {
          float v[8];
          unpack8m(*reinterpret_cast<const uint4*>(dout + ((size_t)b * Lo + 1 + i) * D + c0), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] += v[j];
        }
      }
    } else {
      for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float v[8];
        unpack8m(*reinterpret_cast<const uint4*>(dout + (size_t)b * Lo * D + c0), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += v[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = a[j];
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
      if ((int)threadIdx.x < h)
#pragma unroll
        for (int j = 0; j < 8; ++j) red[threadIdx.x][j] += red[threadIdx.x + h][j];
      __syncthreads();
    }
    if (threadIdx.x < 8) acc_tok[c0 + threadIdx.x] += red[0][threadIdx.x];
    __syncthreads();
  }
  if (acc_pos) {
    for (int t = threadIdx.x; t < Lo; t += blockDim.x) {
      float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int b = 0; b < B; ++b) {
        float v[8];
        unpack8m(*reinterpret_cast<const uint4*>(dout + ((size_t)b * Lo + t) * D + c0), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += v[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc_pos[(size_t)t * D + c0 + j] += s[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// masked-patch MSE.  pred: bf16 rows of PD = p*p*3 values; row of patch (b, l) starts at pred + (b*pred_tokens + pred_off + l)*PD
// (pred_off = 1 skips the cls row of the decoder output in place).  imgs: fp32 NCHW.  One warp per patch.
// ---------------------------------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(256) mae_loss_kernel(const __nv_bfloat16* __restrict__ pred, const float* __restrict__ imgs,
                                                       const float* __restrict__ mask, float* __restrict__ part,
                                                       const float* __restrict__ dloss, __nv_bfloat16* __restrict__ dpred, int B,
                                                       int Hp, int P, int Himg, int pred_tokens, int pred_off, int norm_pix,
                                                       float inv_mask_sum) {
  __shared__ float wsum[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = Hp * Hp, PD = P * P * 3;
  const long long patch = (long long)blockIdx.x * 8 + warp;
  float acc = 0.f;
  if (patch < (long long)B * L) {
    const int b = (int)(patch / L), l = (int)(patch % L);
    const int ph = l / Hp, pw = l % Hp;
    const float mk = mask[patch];
    const __nv_bfloat16* pr = pred + ((size_t)b * pred_tokens + pred_off + l) * PD;
    __nv_bfloat16* dp = BWD ? dpred + ((size_t)b * pred_tokens + pred_off + l) * PD : nullptr;
    if (BWD && l == 0) {   // rows skipped by pred_off (the cls token) carry no loss: their gradient is zero
      __nv_bfloat16* z = dpred + (size_t)b * pred_tokens * PD;
      for (int i = lane * 8; i < pred_off * PD; i += 256) *reinterpret_cast<uint4*>(z + i) = make_uint4(0, 0, 0, 0);
    }
    if (mk != 0.f) {
      // target values in image order (c, pi, qi): coalesced 64-byte runs; element of the patch vector: (pi*P + qi)*3 + c
      float mean = 0.f, rstd = 1.f;
      if (norm_pix) {
        float s = 0.f;
        for (int i = lane; i < PD; i += 32) {
          const int c = i / (P * P), rem = i % (P * P), pi = rem / P, qi = rem % P;
          s += imgs[(((size_t)b * 3 + c) * Himg + ph * P + pi) * Himg + pw * P + qi];
        }
        mean = warp_sum(s) / PD;
        float q = 0.f;
        for (int i = lane; i < PD; i += 32) {
          const int c = i / (P * P), rem = i % (P * P), pi = rem / P, qi = rem % P;
          const float d = imgs[(((size_t)b * 3 + c) * Himg + ph * P + pi) * Himg + pw * P + qi] - mean;
          q = fmaf(d, d, q);
        }
        rstd = rsqrtf(warp_sum(q) / (PD - 1) + 1.e-6f);      // paddle Tensor.var: unbiased
      }
      const float gscale = BWD ? 2.f / PD * mk * inv_mask_sum * (dloss ? dloss[0] : 1.f) : 0.f;
      for (int i = lane; i < PD; i += 32) {
        const int c = i / (P * P), rem = i % (P * P), pi = rem / P, qi = rem % P;
        const float t = (imgs[(((size_t)b * 3 + c) * Himg + ph * P + pi) * Himg + pw * P + qi] - mean) * rstd;
        const int e = (pi * P + qi) * 3 + c;
        const float d = __bfloat162float(pr[e]) - t;
        if (BWD) dp[e] = __float2bfloat16_rn(d * gscale);
        else acc = fmaf(d, d, acc);
      }
      acc = warp_sum(acc) / PD * mk;
    } else if (BWD) {
      for (int i = lane * 8; i < PD; i += 256) *reinterpret_cast<uint4*>(dp + i) = make_uint4(0, 0, 0, 0);
    }
  }
  if (!BWD) {
    if (lane == 0) wsum[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += wsum[w];
      part[blockIdx.x] = s;
    }
  }
}

__global__ void mae_loss_finalize_kernel(const float* __restrict__ part, int n, float inv_mask_sum, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = s * inv_mask_sum;
  }
}

static int mae_blocks(long long n) {
  long long g = (n + 255) / 256;
  long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace pb

using namespace pb;

extern "C" int passl_b200_mae_random_masking(const float* noise, long long* ids_shuffle, long long* ids_restore, float* mask,
                                             int B, int L, int len_keep, void* stream) {
  if (B <= 0 || L <= 0 || L > 1024 || len_keep < 0 || len_keep > L) return PB_ERR_BAD_ARG;
  int LP = 1;
  while (LP < L) LP <<= 1;
  mae_masking_kernel<<<B, 256, LP * 8, (cudaStream_t)stream>>>(noise, ids_shuffle, ids_restore, mask, L, len_keep, LP);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_token_assemble_fwd(const void* src, const long long* ids, const float* pos, const float* tok, void* out,
                                             int B, int Ls, int Lo, int D, int mode, int keep, void* stream) {
  if (B <= 0 || D % 8 || mode < 0 || mode > 2) return PB_ERR_BAD_ARG;
  if (mode != 0 && !ids) return PB_ERR_BAD_ARG;
  token_assemble_fwd_kernel<<<mae_blocks((long long)B * Lo * D / 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), ids, pos, tok, reinterpret_cast<__nv_bfloat16*>(out), B, Ls, Lo, D, mode, keep);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_token_assemble_bwd(const void* dout, const long long* ids, void* dsrc, float* acc_tok, float* acc_pos,
                                             const long long* ids_tok, int B, int Ls, int Lo, int D, int mode, int keep,
                                             void* stream) {
  if (B <= 0 || D % 8 || mode < 0 || mode > 2) return PB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dsrc) {
    token_assemble_bwd_kernel<<<mae_blocks((long long)B * Ls * D / 8), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dout), ids, reinterpret_cast<__nv_bfloat16*>(dsrc), B, Ls, Lo, D, mode, keep);
    PB_LAUNCH_CHECK();
  }
  if (acc_tok || acc_pos) {
    token_param_grad_kernel<<<D / 8, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dout), ids_tok, acc_tok, acc_pos, B, Lo, D,
                                                   mode, keep);
    PB_LAUNCH_CHECK();
  }
  return PB_OK;
}

extern "C" long long passl_b200_mae_loss_workspace_bytes(int B, int L) { return ((long long)B * L / 8 + 2) * 4; }

extern "C" int passl_b200_mae_loss_fwd(const void* pred, const float* imgs, const float* mask, float* loss, int B, int Hp, int P,
                                       int pred_tokens, int pred_off, int norm_pix, float mask_sum, void* workspace, void* stream) {
  if (B <= 0 || Hp <= 0 || (P * P * 3) % 8 || mask_sum <= 0.f) return PB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long patches = (long long)B * Hp * Hp;
  const int nblk = (int)((patches + 7) / 8);
  float* part = reinterpret_cast<float*>(workspace);
  mae_loss_kernel<false><<<nblk, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(pred), imgs, mask, part, nullptr, nullptr, B, Hp,
                                               P, Hp * P, pred_tokens, pred_off, norm_pix, 1.f / mask_sum);
  PB_LAUNCH_CHECK();
  mae_loss_finalize_kernel<<<1, 1024, 0, st>>>(part, nblk, 1.f / mask_sum, loss);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_mae_loss_bwd(const void* pred, const float* imgs, const float* mask, const float* dloss, void* dpred,
                                       int B, int Hp, int P, int pred_tokens, int pred_off, int norm_pix, float mask_sum,
                                       void* stream) {
  if (B <= 0 || Hp <= 0 || (P * P * 3) % 8 || mask_sum <= 0.f) return PB_ERR_BAD_ARG;
  const long long patches = (long long)B * Hp * Hp;
  const int nblk = (int)((patches + 7) / 8);
  mae_loss_kernel<true><<<nblk, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(pred), imgs, mask, nullptr, dloss,
                                                                reinterpret_cast<__nv_bfloat16*>(dpred), B, Hp, P, Hp * P, pred_tokens,
                                                                pred_off, norm_pix, 1.f / mask_sum);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
