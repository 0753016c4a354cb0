// CLIP-specific kernels (passl_v110/modeling/backbones/clip.py:299-338, heads/clip_head.py:27-35):
//
//   token embedding   clip.py:300-303  x = token_embedding(text) + positional_embedding        (gather, int64 ids)
//   EOT pooling       clip.py:307-311  x[i, argmax(text[i])]  ("eot_token is the highest number in each sequence")
//   symmetric CE      clip.py:322-335 + clip_head.py:29-35
//                     C = I_n T_n^T (tensor-core GEMM, fp32 out); s = exp(logit_scale) ON THE DEVICE (no D2H sync);
//                     img_loss = CE_rows(s*C, arange), text_loss = CE_cols(s*C, arange)  (text_logits == img_logits^T);
//                     logit_scale clamped to [-4.6, 4.6] after the forward (clip.py:316-318);
//                     backward: dC = s*dS, dlogit_scale = sum(dS*S), dS = dloss*((softmax_row-1)/n + (softmax_col-1)/n).
// The [n,n] fp32 logits (4 MB at n=1024) stay L2 resident between the three passes; everything here is latency/HBM-bound.
#include <climits>
#include "common.cuh"
#include "host_utils.h"
#include "../../include/passl_b200.h"

namespace pb {

static inline int clip_blocks(long long work, int per_block = 256) {
  long long b = (work + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------------------------------------------------
// token embedding: out[t, :] = table[ids[t], :] + pos[t % L, :]      (8 features per thread, bf16 out)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void embedding_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                     const float* __restrict__ pos, __nv_bfloat16* __restrict__ out, long long T, int L, int D,
                                     int V) {
  const int dv = D / 8;
  const long long total = T * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / dv;
    const int c = (int)(i - t * dv) * 8;
    long long id = ids[t];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    const float4* e = reinterpret_cast<const float4*>(table + id * D + c);
    const float4* p = reinterpret_cast<const float4*>(pos + (size_t)(t % L) * D + c);
    const float4 e0 = e[0], e1 = e[1], p0 = p[0], p1 = p[1];
    uint4 o;
    o.x = pack_bf16x2(e0.x + p0.x, e0.y + p0.y);
    o.y = pack_bf16x2(e0.z + p0.z, e0.w + p0.w);
    o.z = pack_bf16x2(e1.x + p1.x, e1.y + p1.y);
    o.w = pack_bf16x2(e1.z + p1.z, e1.w + p1.w);
    *reinterpret_cast<uint4*>(out + t * D + c) = o;
  }
}

// dtable[ids[t], :] += dout[t, :]   (fp32 reductions in L2: rows collide whenever a token repeats)
__global__ void embedding_bwd_table_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ dout,
                                           float* __restrict__ dtable, long long T, int D, int V) {
  const int dv = D / 8;
  const long long total = T * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / dv;
    const int c = (int)(i - t * dv) * 8;
    long long id = ids[t];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    const uint4 g = *reinterpret_cast<const uint4*>(dout + t * D + c);
    float* dst = dtable + id * D + c;
    float2 f;
    f = unpack_bf16x2(g.x); red_add_f32(dst + 0, f.x); red_add_f32(dst + 1, f.y);
    f = unpack_bf16x2(g.y); red_add_f32(dst + 2, f.x); red_add_f32(dst + 3, f.y);
    f = unpack_bf16x2(g.z); red_add_f32(dst + 4, f.x); red_add_f32(dst + 5, f.y);
    f = unpack_bf16x2(g.w); red_add_f32(dst + 6, f.x); red_add_f32(dst + 7, f.y);
  }
}

// dpos[l, :] += sum_b dout[b*L + l, :]   (one thread per (l, feature pair); deterministic)
__global__ void embedding_bwd_pos_kernel(const __nv_bfloat16* __restrict__ dout, float* __restrict__ dpos, int B, int L, int D) {
  const int dp = D / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * dp) return;
  const int l = i / dp, c = (i - l * dp) * 2;
  float a0 = 0.f, a1 = 0.f;
  for (int b = 0; b < B; ++b) {
    const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + ((size_t)b * L + l) * D + c));
    a0 += f.x;
    a1 += f.y;
  }
  dpos[(size_t)l * D + c] += a0;
  dpos[(size_t)l * D + c + 1] += a1;
}

// ---------------------------------------------------------------------------------------------------------------------
// EOT pooling: idx[i] = first argmax_l ids[i, l];  out[i, :] = x[i*L + idx[i], :]   — one CTA per sample
// ---------------------------------------------------------------------------------------------------------------------
__global__ void eot_gather_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ x,
                                  __nv_bfloat16* __restrict__ out, int* __restrict__ idx, int L, int D) {
  __shared__ long long sv[32];
  __shared__ int si[32];
  __shared__ int s_idx;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long best = LLONG_MIN;
  int bi = 0x7fffffff;
  for (int l = tid; l < L; l += blockDim.x) {
    const long long v = ids[(size_t)b * L + l];
    if (v > best || (v == best && l < bi)) { best = v; bi = l; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (tid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 1; w < nw; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
    s_idx = bi;
    idx[b] = bi;
  }
  __syncthreads();
  const __nv_bfloat16* src = x + ((size_t)b * L + s_idx) * D;
  for (int c = tid * 8; c < D; c += blockDim.x * 8)
    *reinterpret_cast<uint4*>(out + (size_t)b * D + c) = *reinterpret_cast<const uint4*>(src + c);
}

// dx[b*L + l, :] = (l == idx[b]) ? dout[b, :] : 0
__global__ void eot_scatter_kernel(const int* __restrict__ idx, const __nv_bfloat16* __restrict__ dout,
                                   __nv_bfloat16* __restrict__ dx, long long T, int L, int D) {
  const int dv = D / 8;
  const long long total = T * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / dv;
    const int c = (int)(i - t * dv) * 8;
    const long long b = t / L;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if ((int)(t - b * L) == idx[b]) v = *reinterpret_cast<const uint4*>(dout + b * D + c);
    *reinterpret_cast<uint4*>(dx + t * D + c) = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// symmetric cross entropy on C fp32 (cosine similarities, n x n valid inside an ld x ld buffer, ld % 4 == 0 — the GEMM pads
// odd batch sizes); logits S = s*C, s = exp(*logit_scale)
// workspace (floats): [0] s used by this forward, [1..3] pad, [4..4+ld) row lse, [4+ld..4+2ld) col lse, then partials
// ---------------------------------------------------------------------------------------------------------------------
__global__ void clip_ce_rows_kernel(const float* __restrict__ C, const float* __restrict__ logit_scale, float* __restrict__ row_lse,
                                    int n, int ld) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float s = __expf(*logit_scale);
  const float* r = C + (size_t)row * ld;
  float m = -INFINITY, z = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float v = s * r[j];
    if (v > m) { z = z * __expf(m - v); m = v; }
    z += __expf(v - m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), oz = __shfl_xor_sync(0xffffffffu, z, o);
    const float nm = fmaxf(m, om);
    if (nm > -INFINITY) {
      z = z * __expf(m - nm) + oz * __expf(om - nm);
      m = nm;
    }
  }
  if (lane == 0) row_lse[row] = m + __logf(z);
}

// 32 columns per CTA, 32 row-lanes: coalesced over columns, online max/sum over rows
__global__ void clip_ce_cols_kernel(const float* __restrict__ C, const float* __restrict__ logit_scale, float* __restrict__ col_lse,
                                    int n, int ld) {
  __shared__ float sm[32][33], sz[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  const float s = __expf(*logit_scale);
  float m = -INFINITY, z = 0.f;
  if (col < n) {
    for (int i = ty; i < n; i += 32) {
      const float v = s * C[(size_t)i * ld + col];
      if (v > m) { z = z * __expf(m - v); m = v; }
      z += __expf(v - m);
    }
  }
  sm[ty][tx] = m;
  sz[ty][tx] = z;
  __syncthreads();
  if (ty == 0 && col < n) {
    float M = sm[0][tx], Z = sz[0][tx];
    for (int k = 1; k < 32; ++k) {
      const float om = sm[k][tx], oz = sz[k][tx];
      if (oz > 0.f) {
        const float nm = fmaxf(M, om);
        Z = Z * __expf(M - nm) + oz * __expf(om - nm);
        M = nm;
      }
    }
    col_lse[col] = M + __logf(Z);
  }
}

// single CTA: losses = mean(lse - diag); records s and clamps the parameter (clip.py:316-318)
__global__ void clip_ce_finalize_kernel(const float* __restrict__ C, float* __restrict__ logit_scale, float* __restrict__ ws,
                                        float* __restrict__ out3, int n, int ld, int clamp) {
  __shared__ float sa[32], sb[32];
  const float ls = *logit_scale;
  const float s = __expf(ls);
  const float* row_lse = ws + 4;
  const float* col_lse = ws + 4 + ld;
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = s * C[(size_t)i * ld + i];
    a += row_lse[i] - d;
    b += col_lse[i] - d;
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float A = 0.f, Bs = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { A += sa[w]; Bs += sb[w]; }
    out3[0] = A / n;           // img_loss
    out3[1] = Bs / n;          // text_loss
    out3[2] = (A + Bs) / n;    // loss
    ws[0] = s;
    if (clamp) *logit_scale = fminf(fmaxf(ls, -4.6f), 4.6f);
  }
}

// dC = s * dS (bf16, zero in the padding), dS = dloss/n * (exp(S-row_lse) + exp(S-col_lse) - 2*[i==j]);  dlogit_scale += sum(dS*S)
__global__ void clip_ce_bwd_kernel(const float* __restrict__ C, const float* __restrict__ ws, const float* __restrict__ dloss,
                                   __nv_bfloat16* __restrict__ dC, float* __restrict__ part, int n, int ld) {
  __shared__ float sred[32];
  const float s = ws[0];
  const float* row_lse = ws + 4;
  const float* col_lse = ws + 4 + ld;
  const float g = (dloss ? *dloss : 1.f) / n;
  const int lq = ld / 4;
  const long long total = (long long)ld * lq;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / lq), c0 = (int)(i - (long long)r * lq) * 4;
    const long long e = (long long)r * ld + c0;
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < n && c0 < n) {
      const float4 cv = *reinterpret_cast<const float4*>(C + e);
      const float4 cl = *reinterpret_cast<const float4*>(col_lse + c0);
      const float rl = row_lse[r];
      const float v[4] = {cv.x, cv.y, cv.z, cv.w};
      const float l[4] = {cl.x, cl.y, cl.z, cl.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c0 + k < n) {
          const float S = s * v[k];
          float ds = __expf(S - rl) + __expf(S - l[k]);
          if (c0 + k == r) ds -= 2.f;
          ds *= g;
          acc += ds * S;
          d[k] = s * ds;
        }
      }
    }
    uint2 o;
    o.x = pack_bf16x2(d[0], d[1]);
    o.y = pack_bf16x2(d[2], d[3]);
    *reinterpret_cast<uint2*>(dC + e) = o;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sred[w];
    part[blockIdx.x] = t;
  }
}

__global__ void clip_ce_bwd_finalize_kernel(const float* __restrict__ part, int nblk, float* __restrict__ dlogit_scale) {
  __shared__ float sred[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) a += part[i];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sred[w];
    *dlogit_scale += t;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// generic mean cross entropy on materialised logits X [n, m] fp32 with int64 labels (the reference heads' own signature:
// CLIPHead.forward(img_logits, ...), clip_head.py:29-32) — warp per row
// ---------------------------------------------------------------------------------------------------------------------
__global__ void rows_ce_kernel(const float* __restrict__ X, const long long* __restrict__ labels, float* __restrict__ row_loss,
                               float* __restrict__ row_lse, int n, int m) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* r = X + (size_t)row * m;
  float mx = -INFINITY, z = 0.f;
  for (int j = lane; j < m; j += 32) {
    const float v = r[j];
    if (v > mx) { z = z * __expf(mx - v); mx = v; }
    z += __expf(v - mx);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o), oz = __shfl_xor_sync(0xffffffffu, z, o);
    const float nm = fmaxf(mx, om);
    z = z * __expf(mx - nm) + oz * __expf(om - nm);
    mx = nm;
  }
  if (lane == 0) {
    const float lse = mx + __logf(z);
    long long lb = labels[row];
    lb = lb < 0 ? 0 : (lb >= m ? m - 1 : lb);
    row_lse[row] = lse;
    row_loss[row] = lse - r[lb];
  }
}

__global__ void rows_ce_finalize_kernel(const float* __restrict__ row_loss, int n, float* __restrict__ loss) {
  __shared__ float sred[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += row_loss[i];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sred[w];
    *loss = t / n;
  }
}

__global__ void rows_ce_bwd_kernel(const float* __restrict__ X, const long long* __restrict__ labels,
                                   const float* __restrict__ row_lse, const float* __restrict__ dloss, float* __restrict__ dX, int n,
                                   int m) {
  const float g = (dloss ? *dloss : 1.f) / n;
  const long long total = (long long)n * m;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / m), c = (int)(i - (long long)r * m);
    float p = __expf(X[i] - row_lse[r]);
    if (c == (int)labels[r]) p -= 1.f;
    dX[i] = g * p;
  }
}

constexpr int kClipBwdBlocks = 592;

}  // namespace pb

using namespace pb;

extern "C" int passl_b200_embedding_fwd(const long long* ids, const float* table, const float* pos, void* out, long long T, int L,
                                        int D, int V, void* stream) {
  if (T <= 0 || L <= 0 || D <= 0 || D % 8 || V <= 0) return PB_ERR_BAD_ARG;
  embedding_fwd_kernel<<<clip_blocks(T * (D / 8)), 256, 0, (cudaStream_t)stream>>>(ids, table, pos,
                                                                                     reinterpret_cast<__nv_bfloat16*>(out), T, L, D, V);
  PB_LAUNCH_CHECK();
  return 0;
}

extern "C" int passl_b200_embedding_bwd(const long long* ids, const void* dout, float* dtable, float* dpos, long long T, int L, int D,
                                        int V, void* stream) {
  if (T <= 0 || L <= 0 || D <= 0 || D % 8 || V <= 0 || T % L) return PB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(dout);
  if (dtable) {
    embedding_bwd_table_kernel<<<clip_blocks(T * (D / 8)), 256, 0, st>>>(ids, g, dtable, T, D, V);
    PB_LAUNCH_CHECK();
  }
  if (dpos) {
    embedding_bwd_pos_kernel<<<(L * (D / 2) + 127) / 128, 128, 0, st>>>(g, dpos, (int)(T / L), L, D);
    PB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int passl_b200_eot_gather_fwd(const long long* ids, const void* x, void* out, int* idx, int B, int L, int D, void* stream) {
  if (B <= 0 || L <= 0 || D <= 0 || D % 8) return PB_ERR_BAD_ARG;
  eot_gather_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(ids, reinterpret_cast<const __nv_bfloat16*>(x),
                                                         reinterpret_cast<__nv_bfloat16*>(out), idx, L, D);
  PB_LAUNCH_CHECK();
  return 0;
}

extern "C" int passl_b200_eot_gather_bwd(const int* idx, const void* dout, void* dx, int B, int L, int D, void* stream) {
  if (B <= 0 || L <= 0 || D <= 0 || D % 8) return PB_ERR_BAD_ARG;
  const long long T = (long long)B * L;
  eot_scatter_kernel<<<clip_blocks(T * (D / 8)), 256, 0, (cudaStream_t)stream>>>(idx, reinterpret_cast<const __nv_bfloat16*>(dout),
                                                                                   reinterpret_cast<__nv_bfloat16*>(dx), T, L, D);
  PB_LAUNCH_CHECK();
  return 0;
}

extern "C" long long passl_b200_clip_ce_workspace_bytes(int ld) { return (4LL + 2LL * ld + kClipBwdBlocks) * 4; }

extern "C" int passl_b200_clip_ce_fwd(const float* C, float* logit_scale, float* out3, int n, int ld, int clamp, void* workspace,
                                      long long workspace_bytes, void* stream) {
  if (n <= 0 || ld < n || ld % 4 || !C || !logit_scale || !out3) return PB_ERR_BAD_ARG;
  if (workspace_bytes < passl_b200_clip_ce_workspace_bytes(ld)) return PB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = reinterpret_cast<float*>(workspace);
  clip_ce_rows_kernel<<<(n + 7) / 8, 256, 0, st>>>(C, logit_scale, ws + 4, n, ld);
  PB_LAUNCH_CHECK();
  clip_ce_cols_kernel<<<(n + 31) / 32, 1024, 0, st>>>(C, logit_scale, ws + 4 + ld, n, ld);
  PB_LAUNCH_CHECK();
  clip_ce_finalize_kernel<<<1, 256, 0, st>>>(C, logit_scale, ws, out3, n, ld, clamp);
  PB_LAUNCH_CHECK();
  return 0;
}

extern "C" int passl_b200_clip_ce_bwd(const float* C, const float* dloss, void* dC, float* dlogit_scale, int n, int ld,
                                      void* workspace, long long workspace_bytes, void* stream) {
  if (n <= 0 || ld < n || ld % 4 || !C || !dC) return PB_ERR_BAD_ARG;
  if (workspace_bytes < passl_b200_clip_ce_workspace_bytes(ld)) return PB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = reinterpret_cast<float*>(workspace);
  float* part = ws + 4 + 2 * (size_t)ld;
  long long want = ((long long)ld * ld / 4 + 255) / 256;
  const int nblk = (int)(want < kClipBwdBlocks ? want : kClipBwdBlocks);
  clip_ce_bwd_kernel<<<nblk, 256, 0, st>>>(C, ws, dloss, reinterpret_cast<__nv_bfloat16*>(dC), part, n, ld);
  PB_LAUNCH_CHECK();
  if (dlogit_scale) {
    clip_ce_bwd_finalize_kernel<<<1, 256, 0, st>>>(part, nblk, dlogit_scale);
    PB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int passl_b200_rows_ce_fwd(const float* logits, const long long* labels, float* loss, float* row_lse, int n, int m,
                                      void* workspace, long long workspace_bytes, void* stream) {
  if (n <= 0 || m <= 0 || !logits || !labels || !loss || !row_lse) return PB_ERR_BAD_ARG;
  if (workspace_bytes < (long long)n * 4) return PB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* row_loss = reinterpret_cast<float*>(workspace);
  rows_ce_kernel<<<(n + 7) / 8, 256, 0, st>>>(logits, labels, row_loss, row_lse, n, m);
  PB_LAUNCH_CHECK();
  rows_ce_finalize_kernel<<<1, 256, 0, st>>>(row_loss, n, loss);
  PB_LAUNCH_CHECK();
  return 0;
}

extern "C" int passl_b200_rows_ce_bwd(const float* logits, const long long* labels, const float* row_lse, const float* dloss,
                                      float* dlogits, int n, int m, void* stream) {
  if (n <= 0 || m <= 0 || !logits || !labels || !row_lse || !dlogits) return PB_ERR_BAD_ARG;
  rows_ce_bwd_kernel<<<clip_blocks((long long)n * m), 256, 0, (cudaStream_t)stream>>>(logits, labels, row_lse, dloss, dlogits, n, m);
  PB_LAUNCH_CHECK();
  return 0;
}
