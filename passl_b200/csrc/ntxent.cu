// SimCLR NT-Xent + CO2 consistency regulariser over a similarity matrix S (fp32, L2-resident) — forward and dS.
//
// Reference: passl_v110/modeling/heads/simclr_contrastive_head.py:42-102.  With local rows R = [h1; h2] (2n x d) and the
// (all-gathered) columns Z = [h1_all; h2_all] (2m x d), S = R Z^T / T is produced by the tcgen05 GEMM; this kernel
// consumes S row pairs (i, n+i):   u = S[i, :] = [aa_i | ab_i],  v = S[n+i, :] = [ba_i | bb_i],  g = i + rank*n
//   loss_a = LSE(u without k=g)       - u[m+g]         (softmax CE on [ab | aa], aa diagonal masked by -1e9)
//   loss_b = LSE(v without k=m+g)     - v[g]           (softmax CE on [ba | bb], bb diagonal masked)
//   A = softmax(u without {g, m+g}),  B = softmax(v without {g, m+g})      (logit_a / logit_b of the CO2 branch)
//   kl1 = sum B (log B - log A),  kl2 = sum A (log A - log B)              (kl_div(..., 'batchmean') numerators)
//   loss = mean_i(loss_a + loss_b) + 3 * (sum_i kl1 + sum_i kl2) / n,   acc1 = mean(argmax(ab_i) == g)
// Masked entries contribute exactly 0 (exp(-1e9) underflows to 0 in the reference's fp32 as well).
// S is only (2n x 2m) floats (8 MB at n=512, m=512; 32 MB at m=4096): it stays in the 126 MB L2 between the GEMM and the
// three streaming passes here, unlike the MoCo [N, 65537] logits which are never materialised (infonce_tc.cu).
#include "common.cuh"
#include "host_utils.h"
#include "../../include/passl_b200.h"

namespace pb {

struct PairStats {  // per pair i
  float lse_a_full, lse_b_full, lse_a, lse_b, kl1, kl2, loss_ab, correct;
};

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = (l < nw) ? red[l] : (is_max ? -INFINITY : 0.f);
  r = is_max ? warp_max(r) : warp_sum(r);
  return r;  // valid in every thread of every warp
}

// one CTA per pair
// Column layout is described by (c1, c2, period, half): for pair i, c1 = base1 + i is the column of h1_i (masked in u,
// positive of v) and c2 = base2 + i the column of h2_i (positive of u, masked in v); column k belongs to the second view
// ("ab"/"bb" block) iff (k % period) >= half.  Blocked [h1_all; h2_all]: base1 = rank*n, base2 = m + rank*n, period = 2m,
// half = m.  Rank-interleaved (what all_gather([h1; h2]) produces): base1 = rank*2n, base2 = rank*2n + n, period = 2n, half = n.
__global__ void __launch_bounds__(256) ntxent_fwd_kernel(const float* __restrict__ S, int n, int C, int base1, int base2,
                                                         int period, int half, PairStats* __restrict__ stats) {
  __shared__ float red[32];
  __shared__ int arg_red[32];
  const int i = blockIdx.x;
  const int c1 = base1 + i, c2 = base2 + i;
  const float* u = S + (size_t)i * C;
  const float* v = S + (size_t)(n + i) * C;
  // pass 1: maxima (excluding the always-masked own-diagonal entry of each row) and arg-max of the ab block
  float mu = -INFINITY, mv = -INFINITY, best = -INFINITY;
  int besti = -1;
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    float a = u[k], b = v[k];
    if (k != c1) mu = fmaxf(mu, a);
    if (k != c2) mv = fmaxf(mv, b);
    if ((k % period) >= half && (a > best)) { best = a; besti = k; }
  }
  mu = block_reduce(mu, red, true);
  mv = block_reduce(mv, red, true);
  // arg-max over ab (first maximum wins, like paddle argmax/top-1)
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi >= 0 && (besti < 0 || oi < besti))) { best = ob; besti = oi; }
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = best; arg_red[threadIdx.x >> 5] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w)
      if (red[w] > best || (red[w] == best && arg_red[w] >= 0 && (besti < 0 || arg_red[w] < besti))) { best = red[w]; besti = arg_red[w]; }
    arg_red[0] = besti;
  }
  __syncthreads();
  const int amax = arg_red[0];
  // pass 2: exponential sums over the doubly-masked set
  float sa = 0.f, sb = 0.f;
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    if (k == c1 || k == c2) continue;
    sa += __expf(u[k] - mu);
    sb += __expf(v[k] - mv);
  }
  sa = block_reduce(sa, red, false);
  sb = block_reduce(sb, red, false);
  const float lse_a = mu + __logf(sa), lse_b = mv + __logf(sb);
  const float lse_a_full = mu + __logf(sa + __expf(u[c2] - mu));   // adds back the ab positive
  const float lse_b_full = mv + __logf(sb + __expf(v[c1] - mv));   // adds back the ba positive
  // pass 3: KL numerators
  float k1 = 0.f, k2 = 0.f;
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    if (k == c1 || k == c2) continue;
    const float la = u[k] - lse_a, lb = v[k] - lse_b;
    const float r = lb - la;
    k1 += __expf(lb) * r;
    k2 -= __expf(la) * r;
  }
  k1 = block_reduce(k1, red, false);
  k2 = block_reduce(k2, red, false);
  if (threadIdx.x == 0) {
    PairStats s;
    s.lse_a_full = lse_a_full; s.lse_b_full = lse_b_full; s.lse_a = lse_a; s.lse_b = lse_b;
    s.kl1 = k1; s.kl2 = k2;
    s.loss_ab = (lse_a_full - u[c2]) + (lse_b_full - v[c1]);
    s.correct = (amax == c2) ? 1.f : 0.f;
    stats[i] = s;
  }
}

// out[0] = loss, out[1] = acc1 (fraction, like layers.accuracy), out[2] = mean contrast loss, out[3] = co2 = kl1 + kl2
__global__ void ntxent_finalize_kernel(const PairStats* __restrict__ stats, int n, float co2_weight, float* out) {
  __shared__ float red[32];
  float l = 0.f, c = 0.f, k = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { l += stats[i].loss_ab; c += stats[i].correct; k += stats[i].kl1 + stats[i].kl2; }
  l = block_reduce(l, red, false);
  c = block_reduce(c, red, false);
  k = block_reduce(k, red, false);
  if (threadIdx.x == 0) {
    out[0] = l / n + co2_weight * (k / n);
    out[1] = c / n;
    out[2] = l / n;
    out[3] = k / n;
  }
}

// dS (bf16) for both rows of the pair; dloss is a device scalar (upstream gradient) or null.
__global__ void __launch_bounds__(256) ntxent_bwd_kernel(const float* __restrict__ S, const PairStats* __restrict__ stats,
                                                         int n, int C, int base1, int base2, float co2_weight,
                                                         const float* __restrict__ dloss, __nv_bfloat16* __restrict__ dS) {
  const int i = blockIdx.x;
  const int c1 = base1 + i, c2 = base2 + i;
  const float* u = S + (size_t)i * C;
  const float* v = S + (size_t)(n + i) * C;
  __nv_bfloat16* du = dS + (size_t)i * C;
  __nv_bfloat16* dv = dS + (size_t)(n + i) * C;
  const PairStats s = stats[i];
  const float gsc = (dloss ? dloss[0] : 1.f) / n;
  const float w3 = co2_weight;
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const float a = u[k], b = v[k];
    float ga = 0.f, gb = 0.f;
    if (k != c1) ga = __expf(a - s.lse_a_full) - (k == c2 ? 1.f : 0.f);         // d loss_a / d u_k
    if (k != c2) gb = __expf(b - s.lse_b_full) - (k == c1 ? 1.f : 0.f);         // d loss_b / d v_k
    if (k != c1 && k != c2) {
      const float la = a - s.lse_a, lb = b - s.lse_b;
      const float A = __expf(la), B = __expf(lb), r = lb - la;
      ga += w3 * ((A - B) + A * (-r - s.kl2));     // d(kl1 + kl2) / d u_k
      gb += w3 * (B * (r - s.kl1) + (B - A));      // d(kl1 + kl2) / d v_k
    }
    du[k] = __float2bfloat16_rn(ga * gsc);
    dv[k] = __float2bfloat16_rn(gb * gsc);
  }
}

}  // namespace pb

using namespace pb;

extern "C" long long passl_b200_ntxent_workspace_bytes(int n) { return (long long)n * sizeof(PairStats) + 64; }

// S: fp32 [2n, 2m] (= R Z^T / T), Z = all_gather([h1; h2]) i.e. rank-interleaved columns (world = m / n ranks).
// out: fp32[4] = {loss, acc1, contrast, co2}.  workspace keeps the per-pair statistics for the backward call.
extern "C" int passl_b200_ntxent_co2_fwd(const float* S, int n, int m, int rank, float co2_weight, float* out,
                                         void* workspace, long long workspace_bytes, void* stream) {
  if (n <= 0 || m < n || m % n || rank < 0 || (rank + 1) * n > m) return PB_ERR_BAD_ARG;
  if (workspace_bytes < passl_b200_ntxent_workspace_bytes(n)) return PB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  PairStats* stats = reinterpret_cast<PairStats*>(workspace);
  ntxent_fwd_kernel<<<n, 256, 0, st>>>(S, n, 2 * m, rank * 2 * n, rank * 2 * n + n, 2 * n, n, stats);
  PB_LAUNCH_CHECK();
  ntxent_finalize_kernel<<<1, 256, 0, st>>>(stats, n, co2_weight, out);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_ntxent_co2_bwd(const float* S, int n, int m, int rank, float co2_weight, const float* dloss,
                                         void* dS_bf16, const void* workspace, void* stream) {
  if (n <= 0 || m < n) return PB_ERR_BAD_ARG;
  ntxent_bwd_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(S, reinterpret_cast<const PairStats*>(workspace), n, 2 * m,
                                                         rank * 2 * n, rank * 2 * n + n, co2_weight, dloss,
                                                         reinterpret_cast<__nv_bfloat16*>(dS_bf16));
  PB_LAUNCH_CHECK();
  return PB_OK;
}
