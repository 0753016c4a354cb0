// Fused similarity -> tau-scaled softmax -> cross-entropy ("InfoNCE") — fp32 SIMT path.
//
//   logit[i, j] = scale * <A_i, B_j>            j in [0, K)         (never materialised in HBM)
//   optional extra positive column               logit[i, pos] = scale * <A_i, P_i>      (MoCo: [l_pos | l_neg])
//   optional excluded column excl[i]             (SimCLR: masked diagonal, exp(-1e9) == 0 in fp32)
//   loss_i = logsumexp_j logit[i, :] - logit[i, target_i]
//
// Reference math: passl_v110/modeling/architectures/moco.py:178-182 + heads/contrastive_head.py:37-60 (MoCo),
// passl/models/mocov3.py:187-198 (MoCo v3), passl_v110/modeling/backbones/clip.py:320-335 + heads/clip_head.py:27-35.
// This exact-fp32 variant serves BASELINE config C1 (fp32, N=16) and cross-checks the tcgen05 bf16 kernel.
#include "common.cuh"
#include "host_utils.h"
#include "simce_common.cuh"
#include "../../include/passl_b200.h"

namespace pb {

constexpr int CE_TR = 16;    // rows per CTA
constexpr int CE_TK = 128;   // keys per smem tile
constexpr int CE_DC = 128;   // feature chunk
constexpr int CE_KS = CE_DC + 4;  // padded smem row stride (floats)

struct SimCEParams {
  const float* A;
  const void* B;          // float or bf16 [K, D]
  const float* P;         // [N, D] or null
  const long long* label; // [N] or null (required when P == null)
  const int* excl;        // [N] or null
  float scale;
  int N, K, D, splits;
  float* part_m; float* part_l; int* part_cnt;  // [N, splits]
  float* tgt;             // [N]
  // backward
  const float* lse;       // [N]
  const float* grow;      // [N] per-row upstream gradient (already includes 1/N and loss scale)
  float* dA;              // [N, D], atomically accumulated
};

template <typename KT> __device__ __forceinline__ float4 load_key4(const KT* p);
template <> __device__ __forceinline__ float4 load_key4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 load_key4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

// Computes the [CE_TR x CE_TK] logit tile for (row0, key0) into s_tile (smem, row-major [CE_TR][CE_TK]).
template <typename KT>
__device__ __forceinline__ void logits_tile(const SimCEParams& p, int row0, int key0, float* qs, float* ks,
                                            float* s_tile) {
  const int t = threadIdx.x;           // 256 threads
  const int key = t & (CE_TK - 1);
  const int rg = t >> 7;               // 0..1 -> rows rg*8 .. rg*8+7
  float acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.f;
  const KT* Bk = reinterpret_cast<const KT*>(p.B);
  for (int d0 = 0; d0 < p.D; d0 += CE_DC) {
    __syncthreads();
    // q chunk [TR][DC]
    for (int i = t; i < CE_TR * CE_DC / 4; i += 256) {
      int r = i / (CE_DC / 4), c = (i % (CE_DC / 4)) * 4;
      float4 v = make_float4(0, 0, 0, 0);
      if (row0 + r < p.N) v = *reinterpret_cast<const float4*>(p.A + (size_t)(row0 + r) * p.D + d0 + c);
      *reinterpret_cast<float4*>(qs + r * CE_DC + c) = v;
    }
    // key chunk [TK][DC] (padded)
    for (int i = t; i < CE_TK * CE_DC / 4; i += 256) {
      int k = i / (CE_DC / 4), c = (i % (CE_DC / 4)) * 4;
      float4 v = make_float4(0, 0, 0, 0);
      if (key0 + k < p.K) v = load_key4<KT>(Bk + (size_t)(key0 + k) * p.D + d0 + c);
      *reinterpret_cast<float4*>(ks + k * CE_KS + c) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < CE_DC; c += 4) {
      float4 kv = *reinterpret_cast<const float4*>(ks + key * CE_KS + c);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float4 qv = *reinterpret_cast<const float4*>(qs + (rg * 8 + r) * CE_DC + c);
        acc[r] = fmaf(qv.x, kv.x, acc[r]);
        acc[r] = fmaf(qv.y, kv.y, acc[r]);
        acc[r] = fmaf(qv.z, kv.z, acc[r]);
        acc[r] = fmaf(qv.w, kv.w, acc[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) s_tile[(rg * 8 + r) * CE_TK + key] = acc[r] * p.scale;
  __syncthreads();
}

// target logit of each row of the tile -> tg[CE_TR] (smem)
template <typename KT>
__device__ __forceinline__ void target_logits(const SimCEParams& p, int row0, float* tg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const KT* Bk = reinterpret_cast<const KT*>(p.B);
  for (int r = warp; r < CE_TR; r += 8) {
    int row = row0 + r;
    float s = 0.f;
    if (row < p.N) {
      const float* a = p.A + (size_t)row * p.D;
      if (p.P) {
        const float* v = p.P + (size_t)row * p.D;
        for (int d = lane; d < p.D; d += 32) s = fmaf(a[d], v[d], s);
      } else {
        const KT* v = Bk + (size_t)p.label[row] * p.D;
        for (int d = lane; d < p.D; d += 32) s = fmaf(a[d], (float)v[d], s);
      }
    }
    s = warp_sum(s);
    if (lane == 0) tg[r] = s * p.scale;
  }
  __syncthreads();
}

template <typename KT>
__global__ void __launch_bounds__(256) simce_fwd_kernel(const SimCEParams p) {
  extern __shared__ float sm[];
  float* qs = sm;                       // [TR][DC]
  float* ks = qs + CE_TR * CE_DC;       // [TK][KS]
  float* st = ks + CE_TK * CE_KS;       // [TR][TK]
  float* tg = st + CE_TR * CE_TK;       // [TR]
  const int row0 = blockIdx.x * CE_TR;
  const int split = blockIdx.y;
  const int tiles = (p.K + CE_TK - 1) / CE_TK;
  const int t_begin = (int)((long long)split * tiles / p.splits), t_end = (int)((long long)(split + 1) * tiles / p.splits);
  target_logits<KT>(p, row0, tg);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // each warp owns rows warp, warp+8
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  int cnt[2] = {0, 0};
  for (int tile = t_begin; tile < t_end; ++tile) {
    const int key0 = tile * CE_TK;
    logits_tile<KT>(p, row0, key0, qs, ks, st);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = warp + rr * 8, row = row0 + r;
      if (row >= p.N) continue;
      const int ex = p.excl ? p.excl[row] : -1;
      const long long lab = p.P ? -1 : p.label[row];
      float v[4];
      float mx = -INFINITY;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int j = key0 + lane + 32 * u;
        bool ok = (j < p.K) && (j != ex);
        v[u] = ok ? st[r * CE_TK + lane + 32 * u] : -INFINITY;
        mx = fmaxf(mx, v[u]);
        if (ok && j != lab && v[u] > tg[r]) cnt[rr]++;
      }
      mx = warp_max(mx);
      if (mx > -INFINITY) {
        float mn = fmaxf(m[rr], mx);
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) s += (v[u] > -INFINITY) ? __expf(v[u] - mn) : 0.f;
        s = warp_sum(s);
        l[rr] = l[rr] * __expf(m[rr] - mn) + s;
        m[rr] = mn;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int row = row0 + warp + rr * 8;
    if (row >= p.N) continue;
    int c = cnt[rr];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) {
      p.part_m[(size_t)row * p.splits + split] = m[rr];
      p.part_l[(size_t)row * p.splits + split] = l[rr];
      p.part_cnt[(size_t)row * p.splits + split] = c;
      if (split == 0) p.tgt[row] = tg[warp + rr * 8];
    }
  }
}

// dA_i = grow_i * scale * ( sum_j (p_ij - [j == label_i]) B_j  + (p_pos - 1) P_i )
template <typename KT>
__global__ void __launch_bounds__(256) simce_bwd_kernel(const SimCEParams p) {
  extern __shared__ float sm[];
  float* qs = sm;
  float* ks = qs + CE_TR * CE_DC;
  float* st = ks + CE_TK * CE_KS;
  const int row0 = blockIdx.x * CE_TR;
  const int split = blockIdx.y;
  const int tiles = (p.K + CE_TK - 1) / CE_TK;
  const int t_begin = (int)((long long)split * tiles / p.splits), t_end = (int)((long long)(split + 1) * tiles / p.splits);
  const int t = threadIdx.x;
  const KT* Bk = reinterpret_cast<const KT*>(p.B);
  const int nchunks = p.D / CE_DC;  // <= 4
  float acc[4][8];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[c][r] = 0.f;
  const int dcol = t & (CE_DC - 1), rg = t >> 7;
  for (int tile = t_begin; tile < t_end; ++tile) {
    const int key0 = tile * CE_TK;
    logits_tile<KT>(p, row0, key0, qs, ks, st);
    // logits -> gradient coefficients (in place)
    for (int i = t; i < CE_TR * CE_TK; i += 256) {
      int r = i / CE_TK, k = i % CE_TK;
      int row = row0 + r, j = key0 + k;
      float c = 0.f;
      if (row < p.N && j < p.K && !(p.excl && p.excl[row] == j)) {
        float pij = __expf(st[i] - p.lse[row]);
        if (!p.P && p.label[row] == j) pij -= 1.f;
        c = pij * p.grow[row] * p.scale;
      }
      st[i] = c;
    }
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
      if (nchunks > 1 || true) {
        // (re)load key chunk c into ks
        __syncthreads();
        for (int i = t; i < CE_TK * CE_DC / 4; i += 256) {
          int k = i / (CE_DC / 4), cc = (i % (CE_DC / 4)) * 4;
          float4 v = make_float4(0, 0, 0, 0);
          if (key0 + k < p.K) v = load_key4<KT>(Bk + (size_t)(key0 + k) * p.D + c * CE_DC + cc);
          *reinterpret_cast<float4*>(ks + k * CE_KS + cc) = v;
        }
        __syncthreads();
      }
#pragma unroll 4
      for (int k = 0; k < CE_TK; ++k) {
        float kv = ks[k * CE_KS + dcol];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[c][r] = fmaf(st[(rg * 8 + r) * CE_TK + k], kv, acc[c][r]);
      }
    }
    __syncthreads();
  }
  for (int c = 0; c < nchunks; ++c)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int row = row0 + rg * 8 + r;
      if (row < p.N) {
        float v = acc[c][r];
        if (split == 0 && p.P) {
          float ppos = __expf(p.tgt[row] - p.lse[row]);
          v += (ppos - 1.f) * p.grow[row] * p.scale * p.P[(size_t)row * p.D + c * CE_DC + dcol];
        }
        atomicAdd(p.dA + (size_t)row * p.D + c * CE_DC + dcol, v);
      }
    }
}

static int ce_smem_bytes() { return (CE_TR * CE_DC + CE_TK * CE_KS + CE_TR * CE_TK + CE_TR) * 4; }

static int pick_splits(int N, int K) {
  int row_tiles = (N + CE_TR - 1) / CE_TR;
  int tiles = (K + CE_TK - 1) / CE_TK;
  int s = (2 * num_sms() + row_tiles - 1) / row_tiles;
  if (s > tiles) s = tiles;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}

}  // namespace pb

using namespace pb;

extern "C" long long passl_b200_simce_workspace_bytes(int N, int K) {
  int s = pick_splits(N, K);
  return (long long)N * s * 12 + (long long)N * 8 + simce_finalize_scratch_bytes(N) + 256;
}

// Forward.  out_scalars[0..2] = {loss_scale * mean loss, acc1 %, acc5 %}; lse[N] and tgt[N] are saved for backward.
extern "C" int passl_b200_simce_fwd_f32(const float* A, const void* B, int b_is_bf16, const float* P,
                                        const long long* label, const int* excl, float scale, float loss_scale, int N,
                                        int K, int D, float* lse, float* tgt, float* loss_rows, float* out_scalars,
                                        void* workspace, long long workspace_bytes, void* stream) {
  if (N <= 0 || K <= 0 || D <= 0 || D % CE_DC || D > 4 * CE_DC) return PB_ERR_BAD_ARG;
  if (!P && !label) return PB_ERR_BAD_ARG;
  if (workspace_bytes < passl_b200_simce_workspace_bytes(N, K)) return PB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  SimCEParams p{};
  p.A = A; p.B = B; p.P = P; p.label = label; p.excl = excl; p.scale = scale;
  p.N = N; p.K = K; p.D = D; p.splits = pick_splits(N, K);
  char* ws = reinterpret_cast<char*>(workspace);
  p.part_m = reinterpret_cast<float*>(ws); ws += (size_t)N * p.splits * 4;
  p.part_l = reinterpret_cast<float*>(ws); ws += (size_t)N * p.splits * 4;
  p.part_cnt = reinterpret_cast<int*>(ws);
  p.tgt = tgt;
  dim3 grid((N + CE_TR - 1) / CE_TR, p.splits);
  int smem = ce_smem_bytes();
  if (b_is_bf16) {
    PB_CUDA_CHECK(cudaFuncSetAttribute(simce_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    simce_fwd_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(p);
  } else {
    PB_CUDA_CHECK(cudaFuncSetAttribute(simce_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    simce_fwd_kernel<float><<<grid, 256, smem, st>>>(p);
  }
  PB_LAUNCH_CHECK();
  PB_CUDA_CHECK(launch_simce_finalize(p.part_m, p.part_l, p.part_cnt, tgt, N, p.splits, P ? 1 : 0, loss_scale, lse,
                                      loss_rows, out_scalars, reinterpret_cast<float*>(p.part_cnt + (size_t)N * p.splits), st));
  passl_b200_launch_counter_add(1);
  return PB_OK;
}

// Backward w.r.t. A (queries).  dA is overwritten.  dloss: device scalar (upstream grad) or null (=1).
extern "C" int passl_b200_simce_bwd_f32(const float* A, const void* B, int b_is_bf16, const float* P,
                                        const long long* label, const int* excl, float scale, float loss_scale, int N,
                                        int K, int D, const float* lse, const float* tgt, const float* dloss, float* dA,
                                        void* workspace, long long workspace_bytes, void* stream) {
  if (N <= 0 || K <= 0 || D <= 0 || D % CE_DC || D > 4 * CE_DC) return PB_ERR_BAD_ARG;
  if (workspace_bytes < (long long)N * 4) return PB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  SimCEParams p{};
  p.A = A; p.B = B; p.P = P; p.label = label; p.excl = excl; p.scale = scale;
  p.N = N; p.K = K; p.D = D; p.splits = pick_splits(N, K);
  p.lse = lse; p.tgt = const_cast<float*>(tgt); p.dA = dA;
  float* grow = reinterpret_cast<float*>(workspace);
  p.grow = grow;
  fill_rowgrad_kernel<<<(N + 255) / 256, 256, 0, st>>>(grow, dloss, loss_scale / N, N);
  PB_LAUNCH_CHECK();
  PB_CUDA_CHECK(cudaMemsetAsync(dA, 0, (size_t)N * D * 4, st));
  dim3 grid((N + CE_TR - 1) / CE_TR, p.splits);
  int smem = ce_smem_bytes();
  if (b_is_bf16) {
    PB_CUDA_CHECK(cudaFuncSetAttribute(simce_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    simce_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(p);
  } else {
    PB_CUDA_CHECK(cudaFuncSetAttribute(simce_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    simce_bwd_kernel<float><<<grid, 256, smem, st>>>(p);
  }
  PB_LAUNCH_CHECK();
  return PB_OK;
}
